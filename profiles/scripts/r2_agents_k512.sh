# agent profile of the K = 512 forward layer: 2 / 4 A stages, one / two issuing warps, and the barrier protocol alone
cd $GRAFT_REPO_ROOT
export CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/libcgl_prof.so
BASE=$((1|8|32|64|256|512|1024|131072))
PROTO=$((8192|67108864|134217728|268435456|536870912))
for X in 0 2097152 $((2097152|4)) $((PROTO|2097152)) $((PROTO|2097152|4)) $((268435456|536870912|67108864|134217728|2097152)) $((268435456|536870912|67108864|134217728|2097152|4)); do
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_agents.py fwd 512 100 1024 2>&1 | tail -7
done
