set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__cycles_active.avg --clock-control none -k regex:"strip_|reduce_splits|slice_rows_part|direction_stage2" -c 60 --csv --log-file gpurun_out/r02_launches_fused.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget --algorithm gram > gpurun_out/r02_run12_ncu.log 2>&1; echo "ncu rc=$?"
