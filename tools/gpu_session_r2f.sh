mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.txt )
tail -40 gpurun_out/r2f_pytest.txt
timeout 300 python tools/kbench2.py --steps 20 > gpurun_out/r2f_kbench.txt 2>&1; cat gpurun_out/r2f_kbench.txt
