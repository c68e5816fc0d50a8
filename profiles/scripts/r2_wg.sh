cd $GRAFT_REPO_ROOT
for L in "" "cgl-gan_b200/lib/libcgl_prev.so"; do
  echo "== lib ${L:-new}"
  for sh in "1024 100 784" "512 100 1024" "784 200 512" "512 200 256"; do
    CGL_B200_LIB=$L timeout 120 python profiles/adam_bench.py $sh 2>&1 | tail -1
  done
done
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_paths.py -x -q 2>&1 | tail -2
