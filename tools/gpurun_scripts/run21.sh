set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_run21_smoke.log 2>&1; echo "smoke rc=$?"
cat gpurun_out/r02_run21_smoke.log
timeout 900 python tools/parity_report.py > gpurun_out/r02_parity_report.txt 2> gpurun_out/r02_parity_report.err; echo "parity rc=$?"
tail -5 gpurun_out/r02_parity_report.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_run21_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02_run21_tests.log
timeout 1200 python bench.py > gpurun_out/r02_run21_bench.json 2> gpurun_out/r02_run21_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run21_bench.err
timeout 900 python bench.py --impl reference --steps 4 --warmup 3 > gpurun_out/r02_run21_bench_reference.json 2> gpurun_out/r02_run21_bench_reference.err; echo "ref rc=$?"
cat gpurun_out/r02_run21_bench_reference.json | cut -c1-600
