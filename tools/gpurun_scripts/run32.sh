set -x
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:_kernel -s 140 -c 420 --csv --log-file gpurun_out/r02_launches_warm_stream.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget --algorithm stream > gpurun_out/r02_run32_ncu.log 2>&1; echo "ncu rc=$?"
