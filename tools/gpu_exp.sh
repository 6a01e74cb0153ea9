cd $GRAFT_REPO_ROOT
python tools/exp_zone_create.py 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['set_interp_data']['adt_build_ms'], d['set_interp_data']['zone_create_s'], d['set_interp_data']['ms'])"
