set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_fullsize.py -x -q > gpurun_out/pytest_gpu_nccl.log 2>&1; tail -30 gpurun_out/pytest_gpu_nccl.log
