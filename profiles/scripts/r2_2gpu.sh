set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_nccl.py -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err; tail -3 gpurun_out/bench_r2_2gpu.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_r2_2gpu.json"))
print(l["n_gpus"], round(l["value"],1), round(l["ms_per_step"],3), l.get("e2e",{}).get("value"))
for k,v in l["configs"].items(): print(k, v.get("n_gpus"), round(v["value"],1), round(v["ms_per_step"],3), v.get("scaling"))
PY
