cd $GRAFT_REPO_ROOT
python bench.py --steps 5 --warmup 3 2> gpurun_out/b.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print(json.dumps(d['other_configs']['c2b_curvilinear_donor'],indent=1))"
tail -3 gpurun_out/b.err
