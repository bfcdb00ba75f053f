set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_config4.csv python bench.py --workload config4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-fit budget > gpurun_out/r02_run25_ncu.log 2>&1; echo "ncu rc=$?"
