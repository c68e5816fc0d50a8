set -x
cd $GRAFT_REPO_ROOT
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -5 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
for line in open("gpurun_out/bench_${N}gpu.json"):
    if line.startswith("{"):
        l=json.loads(line); print("N=$N value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | grep '^{' | cut -c1-200
