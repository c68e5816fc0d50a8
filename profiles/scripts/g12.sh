set -x
cd $GRAFT_REPO_ROOT
export CGL_TUNE=9
for lib in libcgl_b200.so libcgl_b200_v0.so; do
export CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/$lib
echo "== $lib"
timeout 120 python tests/debug_tc.py 2>&1 | tail -1
for a in "fwd 1024 100 784" "fwd 512 100 1024" "fwd 784 200 512" "fwd 512 200 256" "fwd 784 100 512" "bwd 1024 100 784" "bwd 512 100 1024" "bwd 512 200 256" "bwd 784 100 512"; do timeout 120 python profiles/linear_bench.py $a 1024; done
done
