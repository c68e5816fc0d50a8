set -x
cd $GRAFT_REPO_ROOT
python profiles/adam_bench.py 1024 100 784 1024 > gpurun_out/adam_bench_base.log 2>&1; cat gpurun_out/adam_bench_base.log
python profiles/adam_bench.py 784 200 512 1024 >> gpurun_out/adam_bench_base.log 2>&1; tail -1 gpurun_out/adam_bench_base.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_grouped -c 1 -o gpurun_out/adam_full_r1 -f python profiles/adam_bench.py 1024 100 784 1024 1 > gpurun_out/ncu_adam_full.log 2>&1; tail -3 gpurun_out/ncu_adam_full.log
