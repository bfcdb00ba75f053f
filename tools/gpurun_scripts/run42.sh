set -x
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -s 12 -c 2 -f -o gpurun_out/r02_oz_gemm_full_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget --algorithm stream > gpurun_out/r02_run42_ncu.log 2>&1; echo "ncu rc=$?"
# the per-rank shape of config 3 on 8 GPUs (12 500 samples): the two-level split of both contractions
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -s 12 -c 2 -f -o gpurun_out/r02_oz_gemm_full_rows12500 python bench.py --rows 12500 --steps 2 --warmup 3 --no-cpu-baseline --no-target --e2e-fit budget --algorithm stream > gpurun_out/r02_run42_ncu_rows12500.log 2>&1; echo "ncu rc=$?"
