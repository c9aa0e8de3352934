mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.txt ); tail -6 gpurun_out/r2m_pytest.txt
timeout 300 python tools/kbench2.py --steps 20 > gpurun_out/r2m_kbench.txt 2>&1; cat gpurun_out/r2m_kbench.txt
