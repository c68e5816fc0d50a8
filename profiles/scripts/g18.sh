set -x
cd $GRAFT_REPO_ROOT
CGL_TUNE=41 timeout 120 python tests/debug_tc.py 2>&1 | tail -1
for t in 9 41; do
echo "== TUNE $t"
for a in "fwd 1024 100 784" "fwd 512 100 1024" "fwd 784 200 512" "fwd 512 200 256" "fwd 784 100 512"; do CGL_TUNE=$t timeout 120 python profiles/linear_bench.py $a 1024; done
done
for t in 9 41 9 41; do
CGL_TUNE=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lw$t.json 2> gpurun_out/bench_lw$t.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_lw$t.json"))
print("BENCH tune $t", round(l["ms_per_step"],2), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
done
