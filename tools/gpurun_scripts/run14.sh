set -x
mkdir -p gpurun_out
O=gpurun_out/r02_pair_probe_power.txt
: > $O
for dbg in 0 6 1; do
  LCX_OZ_DEBUG=$dbg timeout 200 python tools/pair_probe.py 100000 10000 100 fp64_split 400 >> $O 2>&1
done
LCX_OZ_CLUSTER=1 timeout 200 python tools/pair_probe.py 100000 10000 100 fp64_split 400 >> $O 2>&1
cat $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -s 12 -c 12 -o gpurun_out/r02_oz_gemm_gram_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget --algorithm gram > gpurun_out/r02_run14_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
