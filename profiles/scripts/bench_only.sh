set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_default.json"))
print(round(l["value"],1), round(l["ms_per_step"],3), l["e2e"]["value"], l["cpu_baseline"]["value"], l["roofline"]["kernel"], l["roofline"]["frac"], l["roofline"]["traffic"], l["roofline_second_kernel"]["kernel"], l["roofline_second_kernel"]["frac"], l["clocks"])
PY
