set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02_run33_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02_run33_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run33_bench.json 2> gpurun_out/r02_run33_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run33_bench.err
