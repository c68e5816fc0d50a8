set -x
cd $GRAFT_REPO_ROOT
for lib in libcgl_b200.so libcgl_b200_v0.so libcgl_b200_v1.so libcgl_b200_v2.so libcgl_b200.so libcgl_b200_v0.so libcgl_b200_v1.so libcgl_b200_v2.so; do
export CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/$lib
echo "== $lib"
python profiles/adam_bench.py 1024 100 784 1024
python profiles/adam_bench.py 784 200 512 1024
python profiles/adam_bench.py 512 100 1024 1024
done
