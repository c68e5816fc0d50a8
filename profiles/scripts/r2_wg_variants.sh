# weight gradient + Adam: TC_WG_UNR (float4 triples in flight per thread in the epilogue; shipped 3) and TC_WG_DEPTH (k-blocks of
# operand loads in flight; shipped 3) as variant libraries
cd $GRAFT_REPO_ROOT
for L in libcgl_b200.so libcgl_wg_unr_2.so libcgl_wg_unr_4.so libcgl_wg_depth_2.so libcgl_wg_depth_4.so libcgl_b200.so; do
  echo "== $L"
  for sh in "1024 100 784" "512 100 1024" "784 200 512" "512 200 256"; do
    CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/$L timeout 120 python profiles/adam_bench.py $sh 2>&1 | tail -1
  done
done
