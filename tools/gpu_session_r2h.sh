mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2h_bench_n1.err; head -c 600 gpurun_out/r2h_bench_n1.json
