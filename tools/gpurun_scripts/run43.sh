set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q --timeout 580 -k "two_level and 4000" > gpurun_out/r02_run43_sanitizer_two_level.log 2>&1; echo "sanitizer rc=$?"
tail -6 gpurun_out/r02_run43_sanitizer_two_level.log
