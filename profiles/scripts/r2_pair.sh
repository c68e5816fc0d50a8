cd $GRAFT_REPO_ROOT
B=$((1|8|32|64|256|512|1024|131072))
SH="fwd:512:100:1024 fwd:512:200:256 bwd:784:100:512 bwd:512:100:1024 fwd:1024:100:784"
echo "== pair, 3 regions"; CGL_TUNE=$((B|8388608)) timeout 300 python profiles/tma_probe.py $SH 2>&1 | tail -5
echo "== pair, lean regions"; CGL_TUNE=$((B|8388608|2097152)) timeout 300 python profiles/tma_probe.py $SH 2>&1 | tail -5
echo "== no pair, lean"; CGL_TUNE=$((B|2097152)) timeout 300 python profiles/tma_probe.py $SH 2>&1 | tail -5
echo "== no pair"; CGL_TUNE=$((B)) timeout 300 python profiles/tma_probe.py $SH 2>&1 | tail -5
echo "== K=100 experiment"
for a in "fwd x 100 128" "fwd p 100 128"; do CGL_TMA_KANY=1 timeout 120 python profiles/tma_repro.py $a 2>&1 | tail -3 | cut -c1-150; done
