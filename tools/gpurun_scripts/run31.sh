set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02_run31_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02_run31_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 500 --csv --log-file gpurun_out/r02_launches_warm_stream.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget --algorithm stream > gpurun_out/r02_run31_ncu.log 2>&1; echo "ncu rc=$?"
