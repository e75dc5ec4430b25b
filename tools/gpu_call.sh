mkdir -p gpurun_out/r2
timeout 180 python tests/gpu_kernels_bw.py > gpurun_out/r2/kernels_bw14.log 2>&1; echo "bw rc=$?"
cat gpurun_out/r2/kernels_bw14.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rhs_jac" --timeout 300 2>&1 | tail -3
