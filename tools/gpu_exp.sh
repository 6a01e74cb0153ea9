cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "synthetic or c4_cluster or mixed_types or distributed" 2>&1 | tail -2
for i in 1 2; do python bench.py --gpus 1 --workload c4 --steps 50 --warmup 10 --quick 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('c4',d['ms_per_step'],d['value'],d['c5_100_iterations']['ms_per_step'],d['gpu_launches'])"; done
