# round 2, session y: the final tree — GPU suite, smoke, both bench arms
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.txt ); tail -3 gpurun_out/r2y_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/r2y_bench_reference_arm.json 2> gpurun_out/r2y_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2y_bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2y_bench_n1.json")); r = d["roofline"]
print("value %.3e e2e %.3e frac %.3f traffic %s hbm frac %.3f" % (d["value"], d["e2e"]["value"], r["frac"], r["traffic"], r["hbm_bound_batch"]["frac"]))
a = d["alt"]
print("one chain", a["emcee_256x2000_one_chain"]["seconds"], a["emcee_256x2000_one_chain"]["seconds_min"], a["emcee_256x2000_one_chain"]["seconds_max"], "nested", a["nested_fit_1000_live_points"]["seconds"], a["nested_fit_1000_live_points"]["converged"])
print("ref", json.load(open("gpurun_out/r2y_bench_reference_arm.json"))["value"])
PY
