set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bench_geometries.py -q --timeout 880 -k "two_level or geometry_8192 or dgemm" > gpurun_out/r02_run44_sanitizer.log 2>&1; echo "sanitizer rc=$?"
tail -6 gpurun_out/r02_run44_sanitizer.log
