set -x
cd $GRAFT_REPO_ROOT
# (each profiled command first runs to completion without ncu)
timeout 200 python profiles/linear_bench.py fwd 1024 100 784 1024 3 > gpurun_out/tma_fwd_plain.log 2>&1; tail -1 gpurun_out/tma_fwd_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_tma -c 1 -o gpurun_out/fwd_tma_r2 -f python profiles/linear_bench.py fwd 1024 100 784 1024 1 > gpurun_out/ncu_tma.log 2>&1; tail -2 gpurun_out/ncu_tma.log
timeout 300 python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; tail -1 gpurun_out/bench_short.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2.csv
