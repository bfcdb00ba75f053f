set -x
mkdir -p gpurun_out
O=gpurun_out/r02_pair_probe_power_elect.txt
: > $O
for dbg in 0 6 1; do
  LCX_OZ_DEBUG=$dbg timeout 200 python tools/pair_probe.py 100000 10000 100 fp64_split 400 >> $O 2>&1
done
timeout 100 python tools/pair_probe.py 100000 10000 100 fp64_split 10 >> $O 2>&1
timeout 100 python tools/pair_probe.py 12500 10000 100 fp64_split 20 >> $O 2>&1
cat $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_run15_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02_run15_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run15_bench.json 2> gpurun_out/r02_run15_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run15_bench.err
timeout 300 python bench.py --workload config4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run15_bench_config4.json 2> gpurun_out/r02_run15_bench_config4.err; echo "bench c4 rc=$?"
