set -x
mkdir -p gpurun_out
bash tools/gpurun_scripts/run12.sh
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gram.py tests/test_gpu_kernels.py tests/test_gpu_properties.py -q --timeout 600 > gpurun_out/r02_run13_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02_run13_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run13_bench.json 2> gpurun_out/r02_run13_bench.err; echo "bench rc=$?"
