set -x
mkdir -p gpurun_out
for i in 1 2; do
for mb in post copy; do
LCX_MAILBOX=$mb timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-target --algorithm gram --e2e-fit budget > gpurun_out/r02_run46_${mb}_$i.json 2> gpurun_out/r02_run46_${mb}_$i.err; echo "bench rc=$?"
done; done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02_run46_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02_run46_tests.log
