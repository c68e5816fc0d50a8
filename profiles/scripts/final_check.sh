# the driver's round-end sequence on one GPU: GPU tests, smoke, default bench, reference arm
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -2 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<PY
import json
for f in ("bench_default","bench_reference"):
    l=json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(l["value"],1), round(l["ms_per_step"],3), l.get("e2e",{}).get("value"), l.get("cpu_baseline",{}).get("value"), l.get("roofline",{}).get("frac") if l.get("roofline") else None)
PY
