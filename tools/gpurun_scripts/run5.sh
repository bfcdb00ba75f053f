set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gram.py -q --timeout 600 -x > gpurun_out/r02_run5_gram_tests.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/r02_run5_gram_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target --algorithm gram > gpurun_out/r02_run5_bench_gram.json 2> gpurun_out/r02_run5_bench_gram.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_run5_bench_gram.err
