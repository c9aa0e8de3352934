# round 2, session d: row-scheduling variants (static+prefetch / claim / claim-ahead+L1 prefetch / claim+register prefetch)
mkdir -p gpurun_out
: > gpurun_out/r2d_kbench.txt
for v in "" var_dyn0 var_dyn3 var_dyn2 var_pair; do
  if [ -z "$v" ]; then unset ISO_B200_LIB; else export ISO_B200_LIB=$PWD/isochrones_b200/lib/$v.so; fi
  timeout 400 python tools/kbench2.py --steps 20 --only posterior,prior,scattered,grid_wide,iso_single,iso_prior,binary,catalog >> gpurun_out/r2d_kbench.txt 2>&1
done
unset ISO_B200_LIB
cat gpurun_out/r2d_kbench.txt
