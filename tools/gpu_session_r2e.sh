# round 2, session e: static scheduling, 48-byte nodes vs EEP-pair records (same box, two repetitions)
mkdir -p gpurun_out
: > gpurun_out/r2e_kbench.txt
for rep in 1 2; do
for v in "" var_pair; do
  if [ -z "$v" ]; then unset ISO_B200_LIB; else export ISO_B200_LIB=$PWD/isochrones_b200/lib/$v.so; fi
  timeout 400 python tools/kbench2.py --steps 20 >> gpurun_out/r2e_kbench.txt 2>&1
done
done
unset ISO_B200_LIB
cat gpurun_out/r2e_kbench.txt
