cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "extract_mesh or golden" 2>&1 | grep -v "NOT YET" | tail -15
