set -x
cd $GRAFT_REPO_ROOT
for a in "fwd 1024 100 784" "fwd 512 100 1024" "fwd 784 200 512" "fwd 512 200 256" "fwd 784 100 512" "bwd 1024 100 784" "bwd 512 100 1024" "bwd 512 200 256" "bwd 784 100 512"; do python profiles/linear_bench.py $a 1024; done 2>&1 | grep -v "^+" > gpurun_out/linear_bench_r1.log; cat gpurun_out/linear_bench_r1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_grouped -c 1 -o gpurun_out/fwd_full_r1 -f python profiles/linear_bench.py fwd 1024 100 784 1024 1 > gpurun_out/ncu_fwd_full.log 2>&1; tail -2 gpurun_out/ncu_fwd_full.log
timeout 300 python profiles/tc_timeline.py 1024 100 784 fwd 1024 2>&1 | tail -11
