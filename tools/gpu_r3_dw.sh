# round 2, session 3: double wall with orders 3 / 5, then the whole GPU suite once more
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_double_wall.py -x -q -m gpu 2>&1 | tail -5
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
