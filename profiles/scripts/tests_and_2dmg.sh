set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_narrow.log 2>&1; tail -12 gpurun_out/pytest_gpu_narrow.log
timeout 300 python bench.py --dataset 2dmg --clients-per-server 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2dmg.json 2> gpurun_out/bench_2dmg.err; tail -2 gpurun_out/bench_2dmg.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_2dmg.json"))
print("BENCH 2dmg", round(l["value"],1), round(l["ms_per_step"],3), {k:round(v["ms_per_round"],3) for k,v in l["kernels"].items()})
PY
