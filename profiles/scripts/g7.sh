set -x
cd $GRAFT_REPO_ROOT
for t in 0 1; do
CGL_TUNE=$t python profiles/adam_bench.py 1024 100 784 1024
CGL_TUNE=$t python profiles/adam_bench.py 784 200 512 1024
CGL_TUNE=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pf$t.json 2> gpurun_out/bench_pf$t.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_pf$t.json"))
print("BENCH tune $t", round(l["ms_per_step"],2), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
done
CGL_TUNE=1 timeout 300 python profiles/tc_timeline.py 1024 100 784 adam 1024 2>&1 | tail -11
