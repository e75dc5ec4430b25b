mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -c 1 -o gpurun_out/r2/prof16 python tests/gpu_probe2.py 148 0 > gpurun_out/r2/ncu16.log 2>&1; echo "ncu rc=$?"
