set -x
mkdir -p gpurun_out/r2
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_c_abi.py tests/test_gpu_parity.py -m gpu -q -s -k "c_driver or multi_device or two_handles" --timeout 300 > gpurun_out/r2/pytest_2gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2/bench_2gpu.json 2> gpurun_out/r2/bench_2gpu.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/r2/bench_2gpu.json | head -c 1500
timeout 900 python bench.py --inlib --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2/bench_2gpu_inlib.json 2> gpurun_out/r2/bench_2gpu_inlib.err; echo "inlib rc=$?"
head -c 600 gpurun_out/r2/bench_2gpu_inlib.json
