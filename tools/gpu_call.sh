mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rhs_jac" > gpurun_out/r2/pytest_k3_51.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest_k3_51.log
timeout 300 python tests/gpu_kernels_bw.py > gpurun_out/r2/kernels_bw_51.log 2>&1; echo "bw rc=$?"
grep "K2\|K3 " gpurun_out/r2/kernels_bw_51.log
