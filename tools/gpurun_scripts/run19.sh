set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --launch-timeout 0 python -m pytest tests/test_gpu_gram.py tests/test_gpu_kernels.py tests/test_gpu_edge_cases.py -q --timeout 1400 -x -k "not 4000x2000 and not adni" > gpurun_out/r02_compute_sanitizer_memcheck.log 2>&1; echo "sanitizer rc=$?"
tail -12 gpurun_out/r02_compute_sanitizer_memcheck.log
