# the generators' Xg pass on a second stream under the clients' D step (CGL_OVERLAP_G=1): e2e arms of the bench (the headline
# arm times per-kernel events and keeps one stream), then the round parity tests with the switch on
cd $GRAFT_REPO_ROOT
for V in 0 1 0 1; do
  CGL_OVERLAP_G=$V timeout 300 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/overlap_$V.json 2> gpurun_out/overlap_$V.err
  python - <<PY
import json
l=json.load(open("gpurun_out/overlap_$V.json"))
print("CGL_OVERLAP_G=$V  headline (one stream, per-kernel events)", round(l["ms_per_step"],3), " e2e", round(l["e2e"]["ms_per_step"],3), " e2e_resident", round(l["e2e_resident"]["ms_per_step"],3))
PY
done
CGL_OVERLAP_G=1 timeout 600 python -m pytest tests/test_gpu_rounds.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
