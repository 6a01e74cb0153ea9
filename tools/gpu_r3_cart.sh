cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_distributed.py -x -q -m gpu -k "cartesian" 2>&1 | tail -4
