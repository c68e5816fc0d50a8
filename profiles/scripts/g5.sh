# usage: bash profiles/scripts/g5.sh <tag>   -- Adam micro-benchmark + instruction counters, short bench
set -x
cd $GRAFT_REPO_ROOT
tag=${1:-x}
python profiles/adam_bench.py 1024 100 784 1024 > gpurun_out/adam_bench_$tag.log 2>&1
python profiles/adam_bench.py 784 200 512 1024 >> gpurun_out/adam_bench_$tag.log 2>&1; cat gpurun_out/adam_bench_$tag.log
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tc_grouped -c 1 python profiles/adam_bench.py 1024 100 784 1024 1 2>&1 | grep -E "inst_executed|duration|issue_active|dram__" 
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_$tag.json"))
print("BENCH", round(l["ms_per_step"],2), l["roofline"]["kernel"], round(l["roofline"]["frac"],3), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
