set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_run17_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r02_run17_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run17_bench.json 2> gpurun_out/r02_run17_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run17_bench.err
