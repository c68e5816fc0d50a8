set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_rounds.py -x -q -k "fegan or fl_round" 2>&1 | tail -25
