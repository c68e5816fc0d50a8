set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2_8gpu.json 2> gpurun_out/bench_r2_8gpu.err; tail -3 gpurun_out/bench_r2_8gpu.err
grep "^{" gpurun_out/bench_r2_8gpu.json | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(l['n_gpus'], round(l['value'],1), round(l['ms_per_step'],3), l.get('e2e',{}).get('value'))
for k,v in l['configs'].items(): print(k, v.get('n_gpus'), round(v['value'],1), round(v['ms_per_step'],3), v.get('scaling'))
"
