# round 2, session z2 (4 GPUs): the bench of the final tree at N = 4 and N = 2
mkdir -p gpurun_out
for n in 4 2; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29720 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2z_bench_n$n.json 2> gpurun_out/r2z_bench_n$n.err
echo "N=$n rc=$?"; tail -c 300 gpurun_out/r2z_bench_n$n.err | tail -2
python - <<PY
import json
d = json.loads(open("gpurun_out/r2z_bench_n$n.json").read().strip().splitlines()[-1])
m = d["multi_gpu"]; ag = m["allgather"]; b = m["binary_1e6_rows_sharded"]; e = m["sharded_ensemble_1M_walkers"]
print("value %.3e e2e %.3e | gather nccl %.4f fused %.4f | binary fused %.4f nccl %.4f none %.4f | catfit %.4f | ens %.4f marginal %.4f overhead %.3f same %s %s" % (
    d["value"], d["e2e"]["value"], ag["ms_per_step_with_gather"], ag["fused_peer_store"]["ms_per_step"], b["ms_per_step_fused_peer_store"],
    b["ms_per_step_nccl_allgather"], b["ms_per_step_without_gather"], m["catalog_10k_stars_fit"]["seconds"], e["ms_per_half_step"],
    e["ms_per_half_step_marginal"], e["run_overhead_ms"], e["copies_identical_across_ranks"], e["identical_to_single_rank_run"]))
PY
done
