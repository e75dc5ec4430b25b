mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2/bench_4gpu.json 2> gpurun_out/r2/bench_4gpu.err; echo "bench4 rc=$?"
head -c 300 gpurun_out/r2/bench_4gpu.json
