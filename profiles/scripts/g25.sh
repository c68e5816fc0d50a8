set -x
cd $GRAFT_REPO_ROOT
for lib in libcgl_b200.so libcgl_b200_v0.so libcgl_b200.so libcgl_b200_v0.so; do
export CGL_B200_LIB=$GRAFT_REPO_ROOT/cgl-gan_b200/lib/$lib
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_d.json"))
print("BENCH $lib", round(l["ms_per_step"],2), {k:round(v["ms_per_round"],2) for k,v in list(l["kernels"].items())[:3]})
PY
done
