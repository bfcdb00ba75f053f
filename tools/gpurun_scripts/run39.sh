set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -q --timeout 500 -k "multi or two_level or gpu_multi or bit_identical or ranks" > gpurun_out/r02_run39_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r02_run39_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_run39_bench_2gpu.json 2> gpurun_out/r02_run39_bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r02_run39_bench_2gpu.err
