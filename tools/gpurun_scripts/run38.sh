set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02_run38_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02_run38_tests.log
for t in 1 0; do
LCX_OZ_TAIL=$t timeout 600 python bench.py --rows 12500 --steps 100 --warmup 5 --no-cpu-baseline --no-target --algorithm stream --e2e-fit budget > gpurun_out/r02_run38_rows12500_tail$t.json 2> gpurun_out/r02_run38_rows12500_tail$t.err; echo "bench rc=$?"
done
for t in 1 0; do
LCX_OZ_TAIL=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target --algorithm stream --e2e-fit budget > gpurun_out/r02_run38_config3_tail$t.json 2> gpurun_out/r02_run38_config3_tail$t.err; echo "bench rc=$?"
done
