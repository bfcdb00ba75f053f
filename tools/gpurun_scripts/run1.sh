set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 120 tools/experiments/bin/mma_mix_probe > gpurun_out/r02_mma_mix_probe.txt 2>&1; echo "probe rc=$?"
timeout 120 tools/experiments/bin/dsmem_feed_probe > gpurun_out/r02_dsmem_feed_probe.txt 2>&1; echo "dsmem rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_run1_tests.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/r02_run1_tests.log
timeout 600 python tools/parity_report.py > gpurun_out/r02_parity_report.txt 2> gpurun_out/r02_parity_report.err; echo "parity rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_run1_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run1_bench.json 2> gpurun_out/r02_run1_bench.err; echo "bench rc=$?"
LCX_OZ_CLUSTER=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target --e2e-fit budget > gpurun_out/r02_run1_bench_cl1.json 2> gpurun_out/r02_run1_bench_cl1.err; echo "bench cl1 rc=$?"
