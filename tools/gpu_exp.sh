cd $GRAFT_REPO_ROOT
python tools/exp_zone_create.py 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --quick --sid-steps 2 --sid-warmup 1 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('adt_build_ms', d['set_interp_data']['adt_build_ms'], 'sid ms', d['set_interp_data']['ms'])"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | grep -v "NOT YET" | tail -2
