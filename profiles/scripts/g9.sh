set -x
cd $GRAFT_REPO_ROOT
for t in 1 3 7; do
echo "== TUNE $t"
for a in "fwd 1024 100 784" "fwd 784 200 512" "bwd 1024 100 784" "bwd 512 100 1024"; do CGL_TUNE=$t python profiles/linear_bench.py $a 1024; done
CGL_TUNE=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_opf$t.json 2> gpurun_out/bench_opf$t.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_opf$t.json"))
print("BENCH tune $t", round(l["ms_per_step"],2), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
done
