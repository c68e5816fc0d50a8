set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2a.log
timeout 900 python bench.py > gpurun_out/bench_default_r2a.json 2> gpurun_out/bench_default_r2a.err; tail -2 gpurun_out/bench_default_r2a.err
timeout 300 python bench.py --dataset 2dmg --clients-per-server 2 --steps 10 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/bench_2dmg_r2a.json 2> gpurun_out/bench_2dmg_r2a.err; tail -2 gpurun_out/bench_2dmg_r2a.err
python - <<PY
import json
for f in ("bench_default_r2a","bench_2dmg_r2a"):
    l=json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(l["value"],1), round(l["ms_per_step"],3), l.get("e2e",{}).get("value"), l.get("cpu_baseline",{}).get("value"), {k:round(v["ms_per_round"],3) for k,v in l.get("kernels",{}).items()})
PY
