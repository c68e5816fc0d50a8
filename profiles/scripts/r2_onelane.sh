# warp-wide barrier waits by one lane + __syncwarp (tune bit 2) against 32 polling lanes
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
PROTO=$((8192|67108864|134217728|268435456|536870912))
for X in 0 2 $PROTO $((PROTO|2)); do
  echo "== extra bits $X"
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_probe.py fwd:1024:100:784 fwd:512:100:1024 fwd:784:200:512 bwd:1024:100:784 2>&1 | grep "bench"
done
CGL_TUNE=$((BASE|2)) timeout 300 python profiles/pair_check.py 2>&1 | grep "check\|worst"
export CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/libcgl_prof.so
for X in 0 2; do
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_agents.py fwd 1024 100 784 2>&1 | tail -6
done
