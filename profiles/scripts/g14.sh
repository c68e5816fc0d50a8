set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1; tail -2 gpurun_out/ncu_traffic.log; wc -l gpurun_out/traffic_r1.csv
nproc; lscpu | grep "Model name"
