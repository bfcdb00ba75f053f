set -x
mkdir -p gpurun_out
timeout 120 tools/experiments/bin/mma_mix_probe > gpurun_out/r02_mma_mix_probe_v2.txt 2>&1; echo "probe rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r02_run2_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r02_run2_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run2_bench.json 2> gpurun_out/r02_run2_bench.err; echo "bench rc=$?"
LCX_OZ_PERSISTENT=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target --e2e-fit budget > gpurun_out/r02_run2_bench_nonpersistent.json 2> gpurun_out/r02_run2_bench_nonpersistent.err; echo "bench np rc=$?"
LCX_OZ_FIXED_KB=10 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target --e2e-fit budget > gpurun_out/r02_run2_bench_fixed10.json 2> gpurun_out/r02_run2_bench_fixed10.err; echo "bench f10 rc=$?"
LCX_OZ_FIXED_KB=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target --e2e-fit budget > gpurun_out/r02_run2_bench_fixed1.json 2> gpurun_out/r02_run2_bench_fixed1.err; echo "bench f1 rc=$?"
timeout 300 python bench.py --workload config4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run2_bench_config4.json 2> gpurun_out/r02_run2_bench_config4.err; echo "bench c4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -s 12 -c 2 -o gpurun_out/r02_oz_gemm_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget > gpurun_out/r02_run2_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
