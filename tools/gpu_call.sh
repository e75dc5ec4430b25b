mkdir -p gpurun_out/r2
timeout 1200 python bench.py > gpurun_out/r2/bench_final.json 2> gpurun_out/r2/bench_final.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2/bench_final.json; tail -3 gpurun_out/r2/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2/bench_ref.json 2> gpurun_out/r2/bench_ref.err; echo "ref rc=$?"
head -c 300 gpurun_out/r2/bench_ref.json
