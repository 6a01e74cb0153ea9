cd $GRAFT_REPO_ROOT
python tools/exp_store_modes.py 2>&1 | tail -3
CX_PLAN_SORT=4 python tools/exp_store_modes.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "transfer or lagrange" 2>&1 | tail -2
