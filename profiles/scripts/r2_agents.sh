# per-agent wait profile of the TMA-fed kernel (library built with -DTCT_PROFILE)
cd $GRAFT_REPO_ROOT
export CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/libcgl_prof.so
BASE=$((1|8|32|64|256|512|1024|131072))
PROTO=$((8192|67108864|134217728|268435456|536870912))
for X in 0 8192 $PROTO $((268435456|536870912|67108864|134217728)) 4096; do
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_agents.py fwd 1024 100 784 2>&1 | tail -7
done
