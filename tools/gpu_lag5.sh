# order-5 Gauss-Jordan A/B (512 vs 1024 threads per receptor) + parity tests of the new transfer features
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lagrange or golden or extract_mesh or rotation or compact" 2>&1 | tail -8
echo "== CX_LAG5=1024"; CX_LAG5=1024 timeout 300 python tests/prof_lagrange.py 5 16 2>&1 | tail -4
echo "== CX_LAG5=1024 parity"; CX_LAG5=1024 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lagrange or golden" 2>&1 | tail -3
echo "== default (512)"; timeout 300 python tests/prof_lagrange.py 5 16 2>&1 | tail -4
