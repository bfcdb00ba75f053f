set -x
mkdir -p gpurun_out
timeout 600 python tools/parity_report.py fp64_split/gram > gpurun_out/r02_parity_report_gram7.txt 2> gpurun_out/r02_parity_report_gram7.err; echo "parity rc=$?"
tail -3 gpurun_out/r02_parity_report_gram7.txt | cut -c1-900
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_run23_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02_run23_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run23_bench.json 2> gpurun_out/r02_run23_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run23_bench.err
timeout 300 python tools/small_configs.py > gpurun_out/r02_run23_small_configs.txt 2>&1; echo "small rc=$?"
cat gpurun_out/r02_run23_small_configs.txt | tail -8
