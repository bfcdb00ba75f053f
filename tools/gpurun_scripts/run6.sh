set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_run6_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r02_run6_tests.log
timeout 900 python bench.py --workload target --steps 20 --warmup 5 --no-cpu-baseline --algorithm gram > gpurun_out/r02_run6_bench_target_gram.json 2> gpurun_out/r02_run6_bench_target_gram.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_run6_bench_target_gram.err
