set -x
mkdir -p gpurun_out
for v in thread inline; do
LCX_HOST_DRAW=$v timeout 200 python tools/e2e_profile.py --per-stage 6 --converge 400 2>&1 | grep "^fit(" > gpurun_out/r02_run53_$v.txt.tmp; cat gpurun_out/r02_run53_$v.txt.tmp >> gpurun_out/r02_run53_$v.txt
done
cat gpurun_out/r02_run53_thread.txt gpurun_out/r02_run53_inline.txt
