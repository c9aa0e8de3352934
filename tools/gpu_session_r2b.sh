# round 2, session b: GPU test suite, kernel timings, bench line, ncu captures of the NODE48 + dynamic-claim kernel
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.txt )
tail -5 gpurun_out/r2b_pytest.txt
timeout 300 python tools/kbench2.py --steps 20 > gpurun_out/r2b_kbench.txt 2>&1; cat gpurun_out/r2b_kbench.txt
timeout 900 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; tail -c 600 gpurun_out/r2b_bench_n1.err; head -c 1500 gpurun_out/r2b_bench_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iso_lnpost_kernel -s 5 -c 1 -f -o gpurun_out/r2b_lnpost_grid_wide python tools/kbench2.py --only grid_wide --steps 2 > gpurun_out/r2b_ncu_gw.log 2>&1; tail -2 gpurun_out/r2b_ncu_gw.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iso_lnpost_kernel -s 5 -c 1 -f -o gpurun_out/r2b_lnpost_posterior python tools/kbench2.py --only posterior --steps 2 > gpurun_out/r2b_ncu_post.log 2>&1; tail -2 gpurun_out/r2b_ncu_post.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:iso_lnpost_kernel -s 5 -c 1 -f -o gpurun_out/r2b_lnpost_binary python tools/kbench2.py --only binary --steps 2 > gpurun_out/r2b_ncu_bin.log 2>&1; tail -2 gpurun_out/r2b_ncu_bin.log
ls -la gpurun_out
