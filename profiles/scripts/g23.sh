set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_head.log 2>&1; tail -3 gpurun_out/pytest_gpu_head.log
for ds in mnist 2dmg; do
extra=""; [ $ds = 2dmg ] && extra="--clients-per-server 2"
timeout 300 python bench.py --dataset $ds $extra --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_head_$ds.json 2> gpurun_out/bench_head_$ds.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_head_$ds.json"))
print("BENCH $ds", round(l["ms_per_step"],3), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
done
