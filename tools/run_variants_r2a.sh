mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/kb_r2a.txt
for v in "" var_dyn1 var_dyn2 var_n64 var_n64_dyn1 var_n48 var_n48_dyn1 var_n48_dyn2 var_n48_dyn1_192 var_n64_dyn1_192 var_pair_dyn1_192; do
  if [ -z "$v" ]; then unset ISO_B200_LIB; else export ISO_B200_LIB=$PWD/isochrones_b200/lib/$v.so; fi
  timeout 300 python tools/kbench2.py --steps 20 >> gpurun_out/kb_r2a.txt 2>&1
done
unset ISO_B200_LIB
cat gpurun_out/kb_r2a.txt
