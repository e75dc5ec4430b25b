mkdir -p gpurun_out/r2
timeout 900 python bench.py --andrews --steps 2 --warmup 3 > gpurun_out/r2/bench_andrews.json 2> gpurun_out/r2/bench_andrews.err; echo "bench rc=$?"
head -c 300 gpurun_out/r2/bench_andrews.json; tail -3 gpurun_out/r2/bench_andrews.err
timeout 600 python bench.py --impl reference --andrews --steps 2 --warmup 1 > gpurun_out/r2/bench_andrews_ref.json 2> gpurun_out/r2/bench_andrews_ref.err; echo "ref rc=$?"
head -c 300 gpurun_out/r2/bench_andrews_ref.json
