# mbarrier wait flavour (variant libraries built with -DTC_WAIT_MODE=1 / 2): full kernel, feed only, barrier protocol only
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
PROTO=$((8192|67108864|134217728|268435456|536870912))
for L in libcgl_b200.so libcgl_wait1.so libcgl_wait2.so; do
for X in 0 8192 $PROTO 4096; do
  echo "== $L extra bits $X"
  CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/$L CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_probe.py fwd:1024:100:784 fwd:512:100:1024 bwd:1024:100:784 2>&1 | grep "bench"
done
done
for L in libcgl_wait1.so libcgl_wait2.so; do
  CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/$L timeout 120 python profiles/adam_bench.py 1024 100 784 2>&1 | tail -1
  CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/$L timeout 120 python profiles/adam_bench.py 784 200 512 2>&1 | tail -1
done
