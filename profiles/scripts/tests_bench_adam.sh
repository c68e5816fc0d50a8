# usage: bash profiles/scripts/g4.sh <tag>   -- GPU parity tests, short bench, Adam micro-benchmark
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-x}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_$tag.json"))
print("BENCH", round(l["ms_per_step"],2), l["roofline"]["kernel"], round(l["roofline"]["frac"],3), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
python profiles/adam_bench.py 1024 100 784 1024 > gpurun_out/adam_bench_$tag.log 2>&1
python profiles/adam_bench.py 784 200 512 1024 >> gpurun_out/adam_bench_$tag.log 2>&1; cat gpurun_out/adam_bench_$tag.log
