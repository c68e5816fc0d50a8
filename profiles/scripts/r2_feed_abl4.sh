# one arrival per B warp on the b_full barrier (default now) against one per thread (tune bit 4): full kernel, feed only, protocol only
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
PROTO=$((8192|67108864|134217728|268435456|536870912))
for X in 0 4 8192 $((8192|4)) $PROTO $((PROTO|4)); do
  echo "== extra bits $X"
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_probe.py fwd:1024:100:784 fwd:512:100:1024 fwd:784:200:512 bwd:1024:100:784 2>&1 | grep "bench"
done
CGL_TUNE=$BASE timeout 300 python profiles/pair_check.py 2>&1 | grep -v "^+"
