# combinations of the opt-in switches on the layers where four A stages fit (K <= 688)
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
for X in 0 $((4|2097152)) $((4|2097152|262144)) $((4|2097152|33554432)) $((4|2097152|262144|33554432)) $((2|2097152)) $((2|2097152|262144|33554432)) 0; do
  echo "== extra bits $X"
  CGL_TUNE=$((BASE|X)) timeout 120 python profiles/tma_probe.py fwd:512:100:1024 fwd:512:200:256 bwd:784:100:512 bwd:1024:100:512 2>&1 | grep "bench"
done
timeout 600 python -m pytest tests/test_gpu_tma.py -x -q -k opt_in 2>&1 | tail -3
