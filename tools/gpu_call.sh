set -x
mkdir -p gpurun_out/r2
timeout 900 python bench.py > gpurun_out/r2/bench_1gpu.json 2> gpurun_out/r2/bench_1gpu.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2/bench_1gpu.json
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:integrate_kernel -c 1 --csv --log-file gpurun_out/r2/traffic_integrate.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
cat gpurun_out/r2/traffic_integrate.csv | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --ncell 2368 > gpurun_out/r2/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
