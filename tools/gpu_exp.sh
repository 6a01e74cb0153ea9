cd $GRAFT_REPO_ROOT
python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(json.dumps(d['set_interp_data']['e2e'],indent=1)); print(d['value'], d['e2e'])"
