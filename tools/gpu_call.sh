mkdir -p gpurun_out/r2
RACG_LIB=rac-2d_b200/libracg_prof.so timeout 300 python tests/gpu_probe2.py 592 0 > gpurun_out/r2/probe31_prof.log 2>&1; echo "probe rc=$?"
tail -5 gpurun_out/r2/probe31_prof.log
timeout 300 python tests/gpu_scan.py gpurun_out/r2/scan31_b60.npz 2 60 0 10000 > gpurun_out/r2/scan31_b60.log 2>&1; echo "scan rc=$?"
tail -2 gpurun_out/r2/scan31_b60.log
timeout 400 python tests/gpu_scan.py gpurun_out/r2/scan31_ext_b60.npz 1 60 10000 70000 > gpurun_out/r2/scan31_ext_b60.log 2>&1; echo "extscan rc=$?"
tail -2 gpurun_out/r2/scan31_ext_b60.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_31.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2/pytest_gpu_31.log
