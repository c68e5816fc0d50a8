set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_tma.py -x -q 2>&1 | tail -5
CGL_PARITY_OUT=$GRAFT_REPO_ROOT/gpurun_out/parity_r2.json timeout 900 python -m pytest tests/test_gpu_parity_report.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline > /dev/null 2> gpurun_out/bench_short.err; tail -1 gpurun_out/bench_short.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic_r2.csv python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1; tail -1 gpurun_out/ncu_traffic.log | cut -c1-200
ls -la gpurun_out/traffic_r2.csv gpurun_out/parity_r2.json
