set -x
mkdir -p gpurun_out
timeout 300 python tools/e2e_profile.py --rows 1000 --vars 50000 --factors 500 --per-stage 2 --converge 0 > gpurun_out/r02_run51_e2e_profile_config4.txt 2>&1; echo "rc=$?"
