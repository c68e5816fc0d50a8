set -x
cd $GRAFT_REPO_ROOT
for t in 105 225 105 225; do
CGL_TUNE=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_ab.json"))
print("BENCH tune $t", round(l["ms_per_step"],2), {k:round(v["ms_per_round"],2) for k,v in list(l["kernels"].items())[:3]})
PY
done
