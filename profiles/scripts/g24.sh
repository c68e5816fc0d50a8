set -x
cd $GRAFT_REPO_ROOT
for t in 41 57 41 57; do
CGL_TUNE=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pf$t.json 2> gpurun_out/bench_pf$t.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_pf$t.json"))
print("BENCH tune $t", round(l["ms_per_step"],2), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
done
