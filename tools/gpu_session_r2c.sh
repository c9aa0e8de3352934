# round 2, session c: layout A/B on catalog / sampler workloads; peer tests (two ranks on one GPU, timeout)
mkdir -p gpurun_out
: > gpurun_out/r2c_kbench.txt
for v in "" var_pair var_n64; do
  if [ -z "$v" ]; then unset ISO_B200_LIB; else export ISO_B200_LIB=$PWD/isochrones_b200/lib/$v.so; fi
  timeout 400 python tools/kbench2.py --steps 20 --only posterior,iso_single,binary,catalog,chains,one_chain >> gpurun_out/r2c_kbench.txt 2>&1
done
unset ISO_B200_LIB
cat gpurun_out/r2c_kbench.txt


