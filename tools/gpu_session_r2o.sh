mkdir -p gpurun_out
timeout 300 python tools/kbench2.py --steps 20 --only posterior,one_chain,chains > gpurun_out/r2o_kbench.txt 2>&1; cat gpurun_out/r2o_kbench.txt
timeout 1200 python bench.py > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2o_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2o_bench_n1.json') if l.startswith('{')][-1])
for k in ('catalog_10k_stars_fit','sharded_ensemble_1M_walkers','emcee_256_walkers_x_1184_chains','emcee_256x2000_one_chain','multinest_calls'):
    f=d['alt'][k]; print(k, {kk:vv for kk,vv in f.items() if kk not in ('config',)})
PY
