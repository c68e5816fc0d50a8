# pipelined A converters (tune bit 262144), with 2 and with 4 A stages (2097152 where the cap allows, 1073741824 timing only)
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
for X in 0 262144 2097152 $((262144|2097152)) $((262144|2097152|1073741824)) $((262144|2097152|33554432)) 0; do
  echo "== extra bits $X"
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_probe.py fwd:1024:100:784 fwd:512:100:1024 fwd:784:200:512 bwd:1024:100:784 bwd:784:100:512 2>&1 | grep "bench"
done
CGL_TUNE=$((BASE|262144|2097152)) timeout 300 python profiles/pair_check.py 2>&1 | grep "check\|worst"
export CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/libcgl_prof.so
for X in 0 $((262144|2097152)); do
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_agents.py fwd 512 100 1024 2>&1 | tail -7
done
