set -x
cd $GRAFT_REPO_ROOT
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
timeout 900 python bench.py --dataset 2dmg --clients-per-server 2 --no-cpu-baseline > gpurun_out/bench_2dmg.json 2> gpurun_out/bench_2dmg.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1; tail -1 gpurun_out/ncu_traffic.log
python - <<PY
import json
for f in ("bench_default","bench_2dmg","bench_reference"):
    l=json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(l["value"],1), round(l["ms_per_step"],3), l.get("e2e",{}).get("value"), l.get("cpu_baseline",{}).get("value"), {k:round(v["ms_per_round"],2) for k,v in l.get("kernels",{}).items()})
PY
