set -x
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r02_run7_bench.json 2> gpurun_out/r02_run7_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_run7_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_gram.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget --algorithm gram > gpurun_out/r02_run7_ncu_gram.log 2>&1; echo "ncu rc=$?"
