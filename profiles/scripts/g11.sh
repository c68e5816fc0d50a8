set -x
cd $GRAFT_REPO_ROOT
export CGL_TUNE=${1:-9}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_persistent -c 1 -o gpurun_out/fwd_persist_full -f python profiles/linear_bench.py fwd 1024 100 784 1024 1 > gpurun_out/ncu_fwd_persist.log 2>&1; tail -2 gpurun_out/ncu_fwd_persist.log
