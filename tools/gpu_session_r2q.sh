# round 2, session q (8 GPUs): the bench at N = 8, 4, 2 as the driver launches it; A/B of the in-kernel completion wait of the
# fused gather (var_peerw0 = separate wait kernel); multi-rank tests with one GPU per rank
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nvidia-smi topo -m > gpurun_out/r2q_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r2q_topo.txt
run() {  # run N tag [env...]
  n=$1; tag=$2; shift 2
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 20 --warmup 5 $EXTRA > gpurun_out/r2q_bench_${tag}.json 2> gpurun_out/r2q_bench_${tag}.err
  echo "N=$n $tag rc=$?"; tail -c 400 gpurun_out/r2q_bench_${tag}.err | tail -3
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2q_bench_${tag}.json").read().strip().splitlines()[-1])
    m = d.get("multi_gpu", {})
    ag = m.get("allgather", {})
    print("  value %.3e e2e %.3e | gather nccl %.4f fused %s | binary %s | catfit %s | ens %s" % (
        d["value"], d["e2e"]["value"], ag.get("ms_per_step_with_gather", -1), ag.get("fused_peer_store", {}).get("ms_per_step"),
        m.get("binary_1e6_rows_sharded", {}).get("ms_per_step_fused_peer_store"), m.get("catalog_10k_stars_fit", {}).get("seconds"),
        m.get("sharded_ensemble_1M_walkers", {}).get("ms_per_half_step")))
except Exception as e:
    print("  parse failed", e)
PY
}
EXTRA=""
run 8 n8
EXTRA="--no-extras --no-cpu-baseline"
run 8 n8_peerw0 ISO_B200_LIB=$PWD/isochrones_b200/lib/var_peerw0.so
run 8 n8_again
EXTRA=""
run 4 n4
run 2 n2
( timeout 240 python -m pytest tests/test_gpu_peer.py tests/test_gpu_ensemble.py -q > gpurun_out/r2q_pytest_8gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest_8gpu.txt ); tail -3 gpurun_out/r2q_pytest_8gpu.txt
