# round-end check on one B200 (run with: gpurun --timeout 2400 -- 'bash tools/gpu_call.sh')
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2/bench_1gpu.json 2> gpurun_out/r2/bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2/bench_ref.json 2> gpurun_out/r2/bench_ref.err; echo "ref rc=$?"
timeout 300 python tests/gpu_kernels_bw.py > gpurun_out/r2/kernels_bw.log 2>&1; echo "bw rc=$?"
