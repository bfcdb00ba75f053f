set -x
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run41_bench.json 2> gpurun_out/r02_run41_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run41_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_run41_bench_reference.json 2> gpurun_out/r02_run41_bench_reference.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_run41_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r02_run41_smoke.log
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -c 1500 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget > gpurun_out/r02_run41_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
