set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_client_fused.py -x -q > gpurun_out/pytest_k1.log 2>&1; tail -30 gpurun_out/pytest_k1.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2b.log 2>&1; tail -15 gpurun_out/pytest_gpu_r2b.log
timeout 300 python bench.py --dataset 2dmg --clients-per-server 2 --steps 10 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/bench_2dmg_r2b.json 2> gpurun_out/bench_2dmg_r2b.err; tail -2 gpurun_out/bench_2dmg_r2b.err
python - <<PY
import json
for f in ("bench_2dmg_r2b",):
    l=json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(l["value"],1), round(l["ms_per_step"],3), l.get("gpu_launches"), {k:(round(v["ms_per_round"],3), round(v["algorithmic_TFLOPps"] or 0,1)) for k,v in l.get("kernels",{}).items()})
PY
