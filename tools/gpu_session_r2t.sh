# round 2, session t (final build, one GPU): full GPU suite, both bench arms, ncu launch list + full captures, memcheck
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.txt ); tail -3 gpurun_out/r2t_pytest.txt
timeout 600 python bench.py --impl reference > gpurun_out/r2t_bench_reference_arm.json 2> gpurun_out/r2t_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2t_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2t_ncu_launches.log 2>&1; tail -2 gpurun_out/r2t_ncu_launches.log | head -c 300
for w in posterior grid_wide binary iso_single; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:iso_lnpost_kernel -s 5 -c 1 -f -o gpurun_out/r2t_lnpost_$w python tools/kbench2.py --only $w --steps 2 > gpurun_out/r2t_ncu_$w.log 2>&1; tail -1 gpurun_out/r2t_ncu_$w.log | head -c 200; echo
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:iso_ensemble_half_kernel -s 4 -c 1 -f -o gpurun_out/r2t_ensemble_half python tools/ens_scale.py --steps 12 > gpurun_out/r2t_ncu_ensemble.log 2>&1; tail -1 gpurun_out/r2t_ncu_ensemble.log | head -c 200; echo
timeout 300 python tools/kbench2.py --steps 20 > gpurun_out/r2t_kbench.txt 2>&1; cat gpurun_out/r2t_kbench.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_golden.py tests/test_gpu_cube.py -q -x > gpurun_out/r2t_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2t_memcheck.txt
ls -la gpurun_out | grep r2t
