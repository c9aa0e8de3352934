# round 2, session n: the bench at N = 2, 4, 8 (one rank per GPU, torchrun as the driver launches it)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2n_bench_n$n.json 2> gpurun_out/r2n_bench_n$n.err
  echo "N=$n rc=$?"; tail -c 600 gpurun_out/r2n_bench_n$n.err | tail -5; head -c 250 gpurun_out/r2n_bench_n$n.json; echo
done
( timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_ensemble.py -q > gpurun_out/r2n_pytest_8gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_8gpu.txt ); tail -3 gpurun_out/r2n_pytest_8gpu.txt
