set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_rounds.py -x -q -k "graph" 2>&1 | tail -15
timeout 300 python profiles/graph_bench.py 2dmg 10 5 100 2>&1 | tail -3
timeout 300 python profiles/graph_bench.py 2dmg 1024 512 30 2>&1 | tail -3
timeout 300 python profiles/graph_bench.py mnist 20 5 50 2>&1 | tail -3
