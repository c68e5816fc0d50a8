set -x
cd $GRAFT_REPO_ROOT
timeout 120 python tests/debug_tc.py 2>&1 | tail -1
for lib in libcgl_b200_v0.so libcgl_b200.so libcgl_b200_v0.so libcgl_b200.so; do
export CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/$lib
echo "== $lib"
python profiles/adam_bench.py 1024 100 784 1024
python profiles/adam_bench.py 784 200 512 1024
python profiles/adam_bench.py 512 100 1024 1024
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_raw.json 2> gpurun_out/bench_raw.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_raw.json"))
print("BENCH $lib", round(l["ms_per_step"],2), {k:round(v["ms_per_round"],2) for k,v in l["kernels"].items()})
PY
done
unset CGL_B200_LIB
timeout 300 python profiles/tc_timeline.py 1024 100 784 adam 1024 2>&1 | tail -11
