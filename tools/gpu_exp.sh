cd $GRAFT_REPO_ROOT
CX_PLAN_SORT=2 CX_TRANSFER_STAGING=0 python tools/exp_store_modes.py 2>&1 | tail -3
