set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2/pytest_gpu_6.log 2>&1; echo "pytest rc=$?"
RACG_LIB=$PWD/rac-2d_b200/libracg_prof.so timeout 300 python tests/gpu_probe2.py 592 0 > gpurun_out/r2/probe6_prof.log 2>&1; echo "probe prof rc=$?"
timeout 400 python tests/gpu_scan.py gpurun_out/r2/scan6_b60.npz 2 60 0 10000 > gpurun_out/r2/scan6_b60.log 2>&1; echo "scan rc=$?"
grep -A7 "^rep 0" gpurun_out/r2/probe6_prof.log
cat gpurun_out/r2/scan6_b60.log
grep -v "^  cell" gpurun_out/r2/pytest_gpu_6.log | tail -60
