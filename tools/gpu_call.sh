mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest_gpu_final2.log
