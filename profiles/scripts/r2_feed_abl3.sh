# A ring depth: 2 A stages in TMEM (default plan) against 4 (2097152: only the regions the cap needs; 1073741824: two hi*hi regions
# at K = 1024, timing only), full kernel and barrier-protocol-only
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
PROTO=$((8192|67108864|134217728|268435456|536870912))
for X in 0 2097152 $((2097152|1073741824)) $PROTO $((PROTO|2097152)) $((PROTO|2097152|1073741824)) $((8192)) $((8192|2097152|1073741824)); do
  echo "== extra bits $X"
  CGL_DEBUG_TMA=1 CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_probe.py fwd:1024:100:784 fwd:512:100:1024 bwd:1024:100:784 2>&1 | grep "bench\|tc_tma" | sort | uniq
done
