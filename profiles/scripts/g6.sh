set -x
cd $GRAFT_REPO_ROOT
tag=${1:-x}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_grouped -c 1 -o gpurun_out/adam_full_$tag -f python profiles/adam_bench.py 1024 100 784 1024 1 > gpurun_out/ncu_adam_full_$tag.log 2>&1; tail -2 gpurun_out/ncu_adam_full_$tag.log
timeout 300 python profiles/tc_timeline.py 1024 100 784 adam 1024 > gpurun_out/tl_adam_$tag.log 2>&1; tail -12 gpurun_out/tl_adam_$tag.log
