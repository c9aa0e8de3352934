mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_ensemble.py tests/test_gpu_cube.py -x -q > gpurun_out/r2q_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.txt ); tail -15 gpurun_out/r2q_pytest.txt
timeout 300 python tools/kbench2.py --steps 20 --only one_chain,chains > gpurun_out/r2q_kbench.txt 2>&1; cat gpurun_out/r2q_kbench.txt
ISO_SAMPLER_QUAD=0 timeout 300 python tools/kbench2.py --steps 20 --only one_chain --tag classic >> gpurun_out/r2q_kbench.txt 2>&1; tail -1 gpurun_out/r2q_kbench.txt
