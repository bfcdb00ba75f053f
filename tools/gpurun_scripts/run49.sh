set -x
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run49_bench.json 2> gpurun_out/r02_run49_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run49_bench.err
