set -x
cd $GRAFT_REPO_ROOT
for a in "fwd x 784 512" "fwd p 784 512" "fwd x 784 512 200" "bwd x 1024 784" "bwd p 1024 784"; do timeout 120 python profiles/tma_repro.py $a 2>&1 | tail -1; done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2d.log 2>&1; tail -5 gpurun_out/pytest_gpu_r2d.log
timeout 900 python bench.py > gpurun_out/bench_default_r2d.json 2> gpurun_out/bench_default_r2d.err; tail -2 gpurun_out/bench_default_r2d.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_default_r2d.json"))
print(round(l["value"],1), round(l["ms_per_step"],3), l.get("e2e",{}).get("value"), l.get("cpu_baseline",{}).get("value"), {k:round(v["ms_per_round"],3) for k,v in l.get("kernels",{}).items()})
PY
