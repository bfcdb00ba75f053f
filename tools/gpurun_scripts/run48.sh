set -x
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run48_bench.json 2> gpurun_out/r02_run48_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run48_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_run48_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/r02_run48_smoke.log
