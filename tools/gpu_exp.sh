cd $GRAFT_REPO_ROOT
for i in 1 2 3; do python tools/exp_store_modes.py 2>&1 | tail -3 | head -1; done
