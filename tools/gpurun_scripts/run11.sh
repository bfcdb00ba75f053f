set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_run11_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r02_run11_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run11_bench.json 2> gpurun_out/r02_run11_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_run11_bench.err
LCX_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target --algorithm gram --e2e-fit budget > gpurun_out/r02_run11_bench_unfused.json 2> gpurun_out/r02_run11_bench_unfused.err; echo "bench rc=$?"
