set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_run30_bench_8gpu.json 2> gpurun_out/r02_run30_bench_8gpu.err; echo "bench8 rc=$?"
tail -3 gpurun_out/r02_run30_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/run_config5.py --max-iter 60 --algorithm auto > gpurun_out/r02_config5_8gpu_auto.json 2> gpurun_out/r02_config5_8gpu_auto.err; echo "auto rc=$?"
cat gpurun_out/r02_config5_8gpu_auto.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/run_config5.py --max-iter 60 --algorithm stream > gpurun_out/r02_config5_8gpu_stream.json 2> gpurun_out/r02_config5_8gpu_stream.err; echo "stream rc=$?"
cat gpurun_out/r02_config5_8gpu_stream.json | cut -c1-1500
