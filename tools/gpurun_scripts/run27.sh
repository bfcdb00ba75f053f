set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_run27_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02_run27_tests.log
timeout 300 python bench.py --workload config4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run27_bench_config4.json 2> gpurun_out/r02_run27_bench_config4.err; echo "bench c4 rc=$?"
