mkdir -p gpurun_out/r2
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:rhs_stream -s 3 -c 1 -o gpurun_out/r2/prof49_k2 -f python tests/gpu_kernels_bw.py > gpurun_out/r2/ncu49_k2.log 2>&1; echo "ncu k2 rc=$?"
ncu -i gpurun_out/r2/prof49_k2.ncu-rep --page raw --csv > gpurun_out/r2/prof49_k2.raw.csv 2>&1
ncu -i gpurun_out/r2/prof49_k2.ncu-rep --page source --csv > gpurun_out/r2/prof49_k2.source.csv 2>&1
ncu -i gpurun_out/r2/prof49_k2.ncu-rep --page details > gpurun_out/r2/prof49_k2.details.txt 2>&1
