# what the feed of the TMA-fed kernel costs without the bytes of one operand (bring-up ablations 67108864 / 134217728)
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
for X in 0 8192 67108864 134217728 $((67108864|134217728)) $((8192|67108864)) $((8192|134217728)) $((8192|67108864|134217728)) 4096; do
  echo "== extra bits $X"
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_probe.py fwd:1024:100:784 fwd:784:200:512 bwd:1024:100:784 2>&1 | grep bench
done
