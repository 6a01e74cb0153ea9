cd $GRAFT_REPO_ROOT
CX_PROFILE=1 python bench.py --gpus 1 --workload c4 --steps 20 --warmup 5 --quick 2>&1 | grep -v "^{" | tail -30
