set -x
cd $GRAFT_REPO_ROOT
CGL_TUNE=105 timeout 120 python tests/debug_tc.py 2>&1 | tail -3
CGL_TUNE=105 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_bk16.log 2>&1; tail -3 gpurun_out/pytest_gpu_bk16.log
for t in 41 105 41 105; do
CGL_TUNE=$t python profiles/adam_bench.py 1024 100 784 1024
CGL_TUNE=$t python profiles/adam_bench.py 784 200 512 1024
CGL_TUNE=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bk.json 2> gpurun_out/bench_bk.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_bk.json"))
print("BENCH tune $t", round(l["ms_per_step"],2), {k:round(v["ms_per_round"],2) for k,v in list(l["kernels"].items())[:3]})
PY
done
