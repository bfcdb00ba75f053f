set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 > gpurun_out/r02_run24_multi_tests.log 2>&1; echo "multi tests rc=$?"
tail -4 gpurun_out/r02_run24_multi_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -s 12 -c 2 -o gpurun_out/r02_oz_gemm_full_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget --algorithm stream > gpurun_out/r02_run24_ncu.log 2>&1; echo "ncu rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget > gpurun_out/r02_run24_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
