set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2c.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2c.log
timeout 900 python bench.py > gpurun_out/bench_default_r2c.json 2> gpurun_out/bench_default_r2c.err; tail -2 gpurun_out/bench_default_r2c.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_default_r2c.json"))
print(round(l["value"],1), round(l["ms_per_step"],3), l.get("e2e",{}).get("value"), l.get("cpu_baseline",{}).get("value"), {k:round(v["ms_per_round"],3) for k,v in l.get("kernels",{}).items()})
PY
