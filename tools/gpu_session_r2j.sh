mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_sampler.py -x -q > gpurun_out/r2j_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.txt ); tail -15 gpurun_out/r2j_pytest.txt
timeout 300 python tools/kbench2.py --steps 20 --only posterior,one_chain,chains > gpurun_out/r2j_kbench.txt 2>&1; cat gpurun_out/r2j_kbench.txt
timeout 1200 python bench.py > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2j_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench_n1.json') if l.startswith('{')][-1])
f=d['alt']['catalog_10k_stars_fit']; print({k:v for k,v in f.items() if k not in ('config','cpu')}); print(f['cpu'])
print(d['alt']['emcee_256_walkers_x_1184_chains']['value'], d['alt']['emcee_256x2000_one_chain']['seconds'])
PY
