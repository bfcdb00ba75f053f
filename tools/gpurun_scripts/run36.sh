set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 100 --warmup 5 --no-target --no-cpu-baseline > gpurun_out/r02_run36_bench_8gpu.json 2> gpurun_out/r02_run36_bench_8gpu.err; echo "bench8 rc=$?"
tail -3 gpurun_out/r02_run36_bench_8gpu.err
