set -x
mkdir -p gpurun_out
O=gpurun_out/r02_run4_pair_probe.txt
: > $O
for dbg in 0 1 2 4 6 7; do
  for cl in 2 1; do
    LCX_OZ_DEBUG=$dbg LCX_OZ_CLUSTER=$cl timeout 120 python tools/pair_probe.py 100000 10000 100 fp64_split 10 >> $O 2>&1
  done
done
for promo in 0 64 128; do
  LCX_OZ_L2PROMO=$promo timeout 120 python tools/pair_probe.py 100000 10000 100 fp64_split 10 >> $O 2>&1
done
for m in 64 128; do
  timeout 120 python tools/pair_probe.py 100000 10000 $m fp64_split 10 >> $O 2>&1
  LCX_OZ_DEBUG=1 timeout 120 python tools/pair_probe.py 100000 10000 $m fp64_split 10 >> $O 2>&1
done
timeout 120 python tools/pair_probe.py 12500 10000 100 fp64_split 20 >> $O 2>&1
LCX_OZ_DEBUG=1 timeout 120 python tools/pair_probe.py 12500 10000 100 fp64_split 20 >> $O 2>&1
cat $O
