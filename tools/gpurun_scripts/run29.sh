set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gram.py -q --timeout 600 > gpurun_out/r02_run29_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02_run29_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run29_bench.json 2> gpurun_out/r02_run29_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run29_bench.err
