mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rhs_jac" > gpurun_out/r2/pytest_k3_40.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest_k3_40.log
timeout 300 python tests/gpu_kernels_bw.py > gpurun_out/r2/kernels_bw_40.log 2>&1; echo "bw rc=$?"
grep "K3 jac (pipe" gpurun_out/r2/kernels_bw_40.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:jac_kernel_pipe -s 3 -c 1 -o gpurun_out/r2/prof40_k3 -f python tests/gpu_kernels_bw.py > gpurun_out/r2/ncu40_k3.log 2>&1; echo "ncu k3 rc=$?"
ncu -i gpurun_out/r2/prof40_k3.ncu-rep --page raw --csv > gpurun_out/r2/prof40_k3.raw.csv 2>&1
ncu -i gpurun_out/r2/prof40_k3.ncu-rep --page source --csv > gpurun_out/r2/prof40_k3.source.csv 2>&1
