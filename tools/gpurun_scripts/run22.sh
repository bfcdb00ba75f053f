set -x
mkdir -p gpurun_out
LCX_SPLIT_DIGITS=7 timeout 300 python - > gpurun_out/r02_run22_adni7.txt 2>&1 <<'PY'
import sys, os
sys.argv=['parity_report.py','fp64_split/gram']
sys.path.insert(0,'tools')
import parity_report as pr
pr.CASES=['adni_l0_f64','adni_l1_f64','readme_demo_f64']
pr.main()
PY
cat gpurun_out/r02_run22_adni7.txt | cut -c1-260
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_run22_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02_run22_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run22_bench.json 2> gpurun_out/r02_run22_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_run22_bench.err
LCX_GRAPH=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-target > gpurun_out/r02_run22_bench_nograph.json 2> gpurun_out/r02_run22_bench_nograph.err; echo "bench rc=$?"
timeout 300 python tools/small_configs.py > gpurun_out/r02_run22_small_configs.txt 2>&1; echo "small rc=$?"
LCX_GRAPH=0 timeout 300 python tools/small_configs.py > gpurun_out/r02_run22_small_configs_nograph.txt 2>&1; echo "small rc=$?"
