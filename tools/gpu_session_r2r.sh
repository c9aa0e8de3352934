# round 2, session r (8 GPUs): partner-gather ensemble + one-fence flag publication; the bench at N = 8, 4, 2
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_ensemble.py -q -x > gpurun_out/r2r_pytest_8gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest_8gpu.txt ); tail -3 gpurun_out/r2r_pytest_8gpu.txt
grep -q "rc=0" gpurun_out/r2r_pytest_8gpu.txt || exit 1
run() {  # run N tag
  n=$1; tag=$2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2r_bench_${tag}.json 2> gpurun_out/r2r_bench_${tag}.err
  echo "N=$n $tag rc=$?"; tail -c 300 gpurun_out/r2r_bench_${tag}.err | tail -2
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2r_bench_${tag}.json").read().strip().splitlines()[-1])
    m = d.get("multi_gpu", {})
    ag = m.get("allgather", {})
    b = m.get("binary_1e6_rows_sharded", {})
    e = m.get("sharded_ensemble_1M_walkers", {})
    print("  value %.3e e2e %.3e | gather nccl %.4f fused %s | binary fused %s nccl %s none %s | catfit %s | ens %s marginal %s overhead %s same %s %s" % (
        d["value"], d["e2e"]["value"], ag.get("ms_per_step_with_gather", -1), ag.get("fused_peer_store", {}).get("ms_per_step"),
        b.get("ms_per_step_fused_peer_store"), b.get("ms_per_step_nccl_allgather"), b.get("ms_per_step_without_gather"),
        m.get("catalog_10k_stars_fit", {}).get("seconds"), e.get("ms_per_half_step"), e.get("ms_per_half_step_marginal"),
        e.get("run_overhead_ms"), e.get("copies_identical_across_ranks"), e.get("identical_to_single_rank_run")))
except Exception as ex:
    print("  parse failed", ex)
PY
}
run 8 n8
run 4 n4
run 2 n2
