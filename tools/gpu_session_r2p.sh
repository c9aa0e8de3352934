mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_nested.py tests/test_gpu_golden.py -x -q > gpurun_out/r2p_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.txt ); tail -30 gpurun_out/r2p_pytest.txt
