set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r02_run8_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r02_run8_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run8_bench.json 2> gpurun_out/r02_run8_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_run8_bench.err
timeout 300 python tools/small_configs.py > gpurun_out/r02_run8_small_configs.txt 2>&1; echo "small rc=$?"
