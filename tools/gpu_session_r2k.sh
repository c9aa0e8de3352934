mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_cube.py -x -q > gpurun_out/r2k_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.txt ); tail -5 gpurun_out/r2k_pytest.txt
timeout 300 python tools/kbench2.py --steps 20 --only posterior,one_chain,chains > gpurun_out/r2k_kbench.txt 2>&1; cat gpurun_out/r2k_kbench.txt
