# which serial chain of the TMA-fed kernel's feed sets its pace: barrier-protocol-only ablations of the A converters (268435456)
# and of the B warps (536870912), with and without MMAs, transfers skipped (67108864 | 134217728)
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
SK=$((67108864|134217728))
for X in 0 268435456 536870912 $((268435456|536870912)) $((SK|268435456|536870912)) $((8192|SK)) $((8192|SK|268435456)) $((8192|SK|536870912)) $((8192|SK|268435456|536870912)); do
  echo "== extra bits $X"
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_probe.py fwd:1024:100:784 bwd:1024:100:784 2>&1 | grep bench
done
