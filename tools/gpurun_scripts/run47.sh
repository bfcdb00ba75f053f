set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 > gpurun_out/r02_run47_multi_tests.log 2>&1; echo "multi tests rc=$?"
tail -3 gpurun_out/r02_run47_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_run47_bench_8gpu.json 2> gpurun_out/r02_run47_bench_8gpu.err; echo "bench8 rc=$?"
tail -3 gpurun_out/r02_run47_bench_8gpu.err
