set -x
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024))
CGL_TUNE=$((BASE|131072)) timeout 300 python profiles/pair_check.py bench > gpurun_out/tma_check.log 2>&1; tail -20 gpurun_out/tma_check.log
CGL_TUNE=$((BASE|131072)) timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -5
