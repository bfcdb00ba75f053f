set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 800 > gpurun_out/r02_run28_multi_tests.log 2>&1; echo "multi tests rc=$?"
tail -30 gpurun_out/r02_run28_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_run28_bench_2gpu.json 2> gpurun_out/r02_run28_bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r02_run28_bench_2gpu.err
