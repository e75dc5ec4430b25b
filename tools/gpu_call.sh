mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/r2/bench_1gpu_final.json 2> gpurun_out/r2/bench_1gpu_final.err; echo "bench rc=$?"
head -c 400 gpurun_out/r2/bench_1gpu_final.json
timeout 600 python bench.py --impl reference > gpurun_out/r2/bench_ref_final.json 2> gpurun_out/r2/bench_ref_final.err; echo "ref rc=$?"
head -c 300 gpurun_out/r2/bench_ref_final.json
