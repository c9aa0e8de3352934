# round 2, session l: the full GPU suite, the KAT tool on a fabricated tree, bench lines, ncu launch list + captures
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.txt ); tail -4 gpurun_out/r2l_pytest.txt
python - > gpurun_out/r2l_kats_fabricated.txt 2>&1 <<'PY'
import os, subprocess, sys, tempfile
sys.path.insert(0, os.getcwd())
from isochrones_b200 import synthetic as syn
from tests.helpers import write_isochrones_tree
root = tempfile.mkdtemp()
bands = ("J", "H", "K", "G", "BP", "RP", "W1", "W2", "W3", "TESS", "Kepler")
bc = syn.make_bc_grid(bands=bands, n_teff=30, n_logg=12, n_feh=8, n_av=7)
write_isochrones_tree(root, "track", syn.make_track_grid(n_mass=40), bc, sidecar=False)
write_isochrones_tree(root, "iso", syn.make_iso_grid(n_age=40), bc, sidecar=True)
r = subprocess.run([sys.executable, "tools/check_mist_kats.py", "--root", root], capture_output=True, text=True)
print(r.stdout[-6000:]); print(r.stderr[-3000:]); print("rc", r.returncode, "(1 = ran to the end; the numbers differ because the tree is synthetic)")
PY
tail -5 gpurun_out/r2l_kats_fabricated.txt
timeout 900 python bench.py --impl reference > gpurun_out/r2l_bench_reference_arm.json 2> gpurun_out/r2l_ref.err; echo "ref rc=$?"
timeout 1200 python bench.py > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2l_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_ncu_launches.log 2>&1; tail -2 gpurun_out/r2l_ncu_launches.log | head -c 300
for w in posterior grid_wide binary iso_single; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:iso_lnpost_kernel -s 5 -c 1 -f -o gpurun_out/r2l_lnpost_$w python tools/kbench2.py --only $w --steps 2 > gpurun_out/r2l_ncu_$w.log 2>&1; tail -1 gpurun_out/r2l_ncu_$w.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iso_sampler_kernel -s 1 -c 1 -f -o gpurun_out/r2l_sampler_one_chain python tools/kbench2.py --only one_chain --steps 2 > gpurun_out/r2l_ncu_sampler.log 2>&1; tail -1 gpurun_out/r2l_ncu_sampler.log
timeout 300 python tools/kbench2.py --steps 20 > gpurun_out/r2l_kbench.txt 2>&1; cat gpurun_out/r2l_kbench.txt
ls -la gpurun_out | head -40
