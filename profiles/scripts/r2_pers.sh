cd $GRAFT_REPO_ROOT
for s in fwd:784:100:512 fwd:528:100:512 fwd:784:200:512 bwd:1024:100:784; do
  timeout 120 python profiles/tma_probe.py $s 2>&1 | grep "^bench\|Error" | head -2
done
timeout 300 python profiles/pair_check.py bench 2>&1 | tail -17
