set -x
cd $GRAFT_REPO_ROOT
tune=${1:-9}
export CGL_TUNE=$tune
timeout 120 python profiles/linear_bench.py fwd 1024 100 784 8 2 || echo "SMALL FAILED rc=$?"
timeout 120 python profiles/linear_bench.py fwd 1024 100 784 1024 || echo "FWD FAILED rc=$?"
timeout 120 python tests/debug_tc.py > gpurun_out/debug_tc_p.log 2>&1; tail -30 gpurun_out/debug_tc_p.log
for a in "fwd 512 100 1024" "fwd 784 200 512" "fwd 512 200 256" "fwd 784 100 512" "bwd 1024 100 784" "bwd 512 100 1024" "bwd 512 200 256" "bwd 784 100 512"; do timeout 120 python profiles/linear_bench.py $a 1024; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_p$tune.log 2>&1; tail -5 gpurun_out/pytest_gpu_p$tune.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p$tune.json 2> gpurun_out/bench_p$tune.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_p$tune.json"))
print("BENCH tune $tune", round(l["ms_per_step"],2), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
