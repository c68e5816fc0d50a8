cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024))
SH="fwd:1024:100:784 fwd:512:100:1024 fwd:784:200:512 bwd:784:100:512 bwd:1024:100:784"
for T in $((BASE|131072)) $((BASE|131072|1048576)) $((BASE|131072|4096)) $((BASE|131072|8192)); do
  echo "== CGL_TUNE=$T"
  CGL_TUNE=$T timeout 200 python profiles/tma_probe.py $SH 2>&1 | tail -6
done > gpurun_out/tma_elect.log 2>&1
cat gpurun_out/tma_elect.log
CGL_TUNE=$((BASE|131072)) timeout 300 python profiles/pair_check.py 2>&1 | tail -8
