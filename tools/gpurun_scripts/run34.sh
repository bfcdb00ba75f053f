set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run34_bench_pdl1.json 2> gpurun_out/r02_run34_bench_pdl1.err; echo "bench rc=$?"
LCX_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run34_bench_pdl0.json 2> gpurun_out/r02_run34_bench_pdl0.err; echo "bench rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02_run34_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02_run34_tests.log
