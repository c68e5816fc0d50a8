set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --algo flgan --steps 5 --warmup 3 --cpu-sample-clients 8 > gpurun_out/bench_flgan.json 2> gpurun_out/bench_flgan.err; tail -3 gpurun_out/bench_flgan.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_flgan.json"))
print("FLGAN", round(l["value"],1), round(l["ms_per_step"],3), l["e2e"]["value"], l.get("cpu_baseline",{}).get("value"), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
