set -x
mkdir -p gpurun_out/r2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jac_kernel4 -c 1 -o gpurun_out/r2/prof_k3 python tests/gpu_kernels_bw.py > gpurun_out/r2/ncu_k3.log 2>&1; echo "ncu rc=$?"
timeout 1500 python bench.py --network rate12-withGrain --ncell 2000 --steps 2 --warmup 2 --cpu-cells 32 --kernel-ncell 37888 > gpurun_out/r2/bench_rate12.json 2> gpurun_out/r2/bench_rate12.err; echo "rate12 rc=$?"
head -c 300 gpurun_out/r2/bench_rate12.json; tail -3 gpurun_out/r2/bench_rate12.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "rate12" --timeout 500 > gpurun_out/r2/pytest_rate12.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2/pytest_rate12.log
