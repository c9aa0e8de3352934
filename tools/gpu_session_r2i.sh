mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2i_bench_n2.err; head -c 400 gpurun_out/r2i_bench_n2.json
( timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_ensemble.py -x -q > gpurun_out/r2i_pytest_2gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_2gpu.txt ); tail -5 gpurun_out/r2i_pytest_2gpu.txt
