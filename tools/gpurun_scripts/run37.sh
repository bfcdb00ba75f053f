set -x
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run37_bench_$i.json 2> gpurun_out/r02_run37_bench_$i.err; echo "bench rc=$?"
done
tail -5 gpurun_out/r02_run37_bench_1.err
