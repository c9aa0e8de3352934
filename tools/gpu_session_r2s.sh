# round 2, session s (8 GPUs): sharded ensemble only, sector-wide partner gathers
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_ensemble.py -q -x > gpurun_out/r2s_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.txt ); tail -3 gpurun_out/r2s_pytest.txt
grep -q "rc=0" gpurun_out/r2s_pytest.txt || exit 1
for n in 1 2 4 8; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) tools/ens_scale.py 2> gpurun_out/r2s_ens_n$n.err | tail -1 | tee -a gpurun_out/r2s_ens_scale.jsonl
done
