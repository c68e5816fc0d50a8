# weight gradient + Adam: what the epilogue stream costs without the operand loads (4096) / the MMAs (8192) / both
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
for X in 0 4096 8192 $((4096|8192)) 0; do
  echo "== extra bits $X"
  for sh in "1024 100 784" "512 100 1024" "784 200 512" "512 200 256"; do
    CGL_TUNE=$((BASE|X)) timeout 120 python profiles/adam_bench.py $sh 2>&1 | tail -1
  done
done
