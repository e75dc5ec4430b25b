set -x
mkdir -p gpurun_out/r2
RACG_LIB=$PWD/rac-2d_b200/libracg_prof.so timeout 300 python tests/gpu_probe2.py 592 0 > gpurun_out/r2/probe5_prof.log 2>&1; echo "probe prof rc=$?"
grep -A7 "^rep 0" gpurun_out/r2/probe5_prof.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -c 1 -o gpurun_out/r2/prof5 python tests/gpu_probe2.py 148 0 > gpurun_out/r2/ncu5.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2/
