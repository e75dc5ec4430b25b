mkdir -p gpurun_out/r2
RACG_LIB=$PWD/rac-2d_b200/libracg_prof.so timeout 300 python tests/gpu_probe2.py 592 8 > gpurun_out/r2/probe15_prof.log 2>&1; echo "probe rc=$?"
grep -A16 "^rep 0" gpurun_out/r2/probe15_prof.log
timeout 400 python tests/gpu_scan.py gpurun_out/r2/scan15_b60.npz 2 60 0 10000 > gpurun_out/r2/scan15_b60.log 2>&1; cat gpurun_out/r2/scan15_b60.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
