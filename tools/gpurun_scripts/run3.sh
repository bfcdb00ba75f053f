set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 > gpurun_out/r02_run3_multi_tests.log 2>&1; echo "multi tests rc=$?"
tail -15 gpurun_out/r02_run3_multi_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_run3_bench_2gpu.json 2> gpurun_out/r02_run3_bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r02_run3_bench_2gpu.err
LCX_OZ_PERSISTENT=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-target --e2e-fit budget > gpurun_out/r02_run3_bench_2gpu_np.json 2> gpurun_out/r02_run3_bench_2gpu_np.err; echo "bench2 np rc=$?"
