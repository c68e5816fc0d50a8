# A/B of the truncated-hi batch operand (tune bit 33554432) in the TMA-fed kernel: error against float64, per-layer times, the round
set -x
cd $GRAFT_REPO_ROOT
BASE=$((1|8|32|64|256|512|1024|131072))
for T in $BASE $((BASE|33554432)); do
  CGL_TUNE=$T timeout 300 python profiles/pair_check.py bench > gpurun_out/trunc_check_$T.log 2>&1; cat gpurun_out/trunc_check_$T.log | grep -v "^+"
done
CGL_TUNE=$((BASE|33554432)) timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tma.py tests/test_gpu_rounds.py -x -q 2>&1 | tail -3
for T in $BASE $((BASE|33554432)); do
  CGL_TUNE=$T timeout 300 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/trunc_bench_$T.json 2> gpurun_out/trunc_bench_$T.err
  python - <<PY
import json
l=json.load(open("gpurun_out/trunc_bench_$T.json"))
print("TUNE $T", round(l["ms_per_step"],3), {k:round(v["ms_per_round"],3) for k,v in l.get("kernels",{}).items()})
PY
done
