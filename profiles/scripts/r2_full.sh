set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2e.log 2>&1; tail -4 gpurun_out/pytest_gpu_r2e.log
timeout 900 python bench.py > gpurun_out/bench_default_r2e.json 2> gpurun_out/bench_default_r2e.err; tail -2 gpurun_out/bench_default_r2e.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_default_r2e.json"))
print(round(l["value"],1), round(l["ms_per_step"],3), l.get("e2e",{}).get("value"), l.get("cpu_baseline",{}).get("value"), {k:round(v["ms_per_round"],3) for k,v in l.get("kernels",{}).items()})
for k,v in l["configs"].items(): print(k, round(v["value"],1), round(v["ms_per_step"],3))
PY
