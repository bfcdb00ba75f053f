set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 500 > gpurun_out/r02_run52_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02_run52_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run52_bench.json 2> gpurun_out/r02_run52_bench.err; echo "bench rc=$?"
