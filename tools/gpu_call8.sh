mkdir -p gpurun_out/r2
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2/bench_8gpu.json 2> gpurun_out/r2/bench_8gpu.err; echo "bench8 rc=$?"
tail -c 1200 gpurun_out/r2/bench_8gpu.json | head -c 1200
