mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_cube.py tests/test_gpu_sampler.py tests/test_gpu_oracle.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2g_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.txt )
tail -40 gpurun_out/r2g_pytest.txt
