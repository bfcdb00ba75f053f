set -x
mkdir -p gpurun_out
timeout 300 python bench.py --workload config4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run50_bench_config4.json 2> gpurun_out/r02_run50_bench_config4.err; echo "bench c4 rc=$?"
tail -3 gpurun_out/r02_run50_bench_config4.err
LCX_OZ_TAIL=0 timeout 300 python bench.py --workload config4 --steps 20 --warmup 5 --no-cpu-baseline --e2e-fit budget > gpurun_out/r02_run50_bench_config4_notail.json 2> gpurun_out/r02_run50_bench_config4_notail.err; echo "bench c4 rc=$?"
