mkdir -p gpurun_out
timeout 60 python tools/quad_debug.py 8 > gpurun_out/r2s_dbg.txt 2>&1; echo "dbg rc=$?"; tail -3 gpurun_out/r2s_dbg.txt
( timeout 600 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_ensemble.py -x -q > gpurun_out/r2s_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.txt ); tail -8 gpurun_out/r2s_pytest.txt
timeout 120 python tools/kbench2.py --steps 20 --only one_chain,chains > gpurun_out/r2s_kbench.txt 2>&1; cat gpurun_out/r2s_kbench.txt
ISO_SAMPLER_QUAD=0 timeout 120 python tools/kbench2.py --steps 20 --only one_chain --tag classic >> gpurun_out/r2s_kbench.txt 2>&1; tail -1 gpurun_out/r2s_kbench.txt
