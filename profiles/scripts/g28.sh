set -x
cd $GRAFT_REPO_ROOT
for t in 105 104 105 104; do
echo "== tune $t"
CGL_TUNE=$t python profiles/adam_bench.py 1024 100 784 1024
CGL_TUNE=$t python profiles/adam_bench.py 784 200 512 1024
done
CGL_TUNE=105 timeout 300 python profiles/tc_timeline.py 1024 100 784 adam 1024 2>&1 | tail -11
