# round 2: full-size parity gate alone (prints the oracle's build/search times), then the rest of the GPU suite
cd $GRAFT_REPO_ROOT
nproc
timeout 1700 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -s 2>&1 | tail -40
