mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 > gpurun_out/r2/pytest_gpu_13.log 2>&1; echo "pytest rc=$?"
grep -v "^  cell" gpurun_out/r2/pytest_gpu_13.log | tail -30 | cut -c1-220
