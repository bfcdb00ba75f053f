set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py tests/test_gpu_properties.py -q --timeout 600 > gpurun_out/r02_run18_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02_run18_tests.log
