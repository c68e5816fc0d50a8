set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_client_fused.py tests/test_gpu_paths.py tests/test_gpu_kernels.py -x -q > gpurun_out/pytest_k1.log 2>&1; tail -15 gpurun_out/pytest_k1.log
python profiles/k1_bench.py 1024 20
CGL_K1=0 python profiles/k1_bench.py 1024 20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:client_step_fused -c 1 -o gpurun_out/k1_mma -f python profiles/k1_bench.py 296 1 > gpurun_out/ncu_k1.log 2>&1; tail -2 gpurun_out/ncu_k1.log
