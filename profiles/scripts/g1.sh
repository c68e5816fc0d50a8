set -x
cd $GRAFT_REPO_ROOT
CGL_TUNE=3 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_20.log 2>&1; tail -3 gpurun_out/pytest_gpu_20.log
for t in 0 1 2 3; do
  CGL_TUNE=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tune$t.json 2> gpurun_out/bench_tune$t.err
  python - <<PY
import json
l=json.load(open("gpurun_out/bench_tune$t.json"))
print("TUNE $t", l["ms_per_step"], l["roofline"]["kernel"], l["roofline"]["frac"], {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
done
CGL_TUNE=1 timeout 300 python profiles/tc_timeline.py 1024 100 784 adam 1024 > gpurun_out/tl_adam_t1.log 2>&1; tail -12 gpurun_out/tl_adam_t1.log
