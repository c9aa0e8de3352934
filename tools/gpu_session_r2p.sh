# round 2, session p: full GPU suite on the in-kernel peer wait + multi-chunk star-sequential kernels; A/B of the latter
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.txt ); tail -6 gpurun_out/r2p_pytest.txt
for lib in "" isochrones_b200/lib/var_seq1.so; do
  ISO_B200_LIB=$lib timeout 300 python tools/kbench2.py --only binary,binary8,binary11 --steps 20 >> gpurun_out/r2p_kbench_seq.txt 2>&1
done
cat gpurun_out/r2p_kbench_seq.txt
timeout 300 python tools/kbench2.py --steps 20 > gpurun_out/r2p_kbench.txt 2>&1; cat gpurun_out/r2p_kbench.txt
