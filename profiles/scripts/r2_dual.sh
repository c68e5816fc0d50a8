# two issuing warps that take the k-blocks in turn (default) against one (tune bit 4)
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
CGL_TUNE=$BASE timeout 120 python profiles/pair_check.py 2>&1 | grep "check\|worst\|rror"
for X in 4 0 $((2097152)) $((2097152|262144)) 4 0; do
  echo "== extra bits $X"
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_probe.py fwd:1024:100:784 fwd:512:100:1024 fwd:784:200:512 bwd:1024:100:784 bwd:784:100:512 2>&1 | grep "bench"
done
export CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/libcgl_prof.so
for X in 4 0; do
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_agents.py fwd 1024 100 784 2>&1 | tail -7
done
