cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "NOT YET" | tail -3
CX_PROFILE_ZONE=1 CX_PROFILE=1 python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick 2>&1 | grep -v "^{" | tail -6 | cut -c1-250
CX_PROFILE_ZONE=1 CX_PROFILE=1 python bench.py --gpus 1 --workload c2n --steps 5 --warmup 3 --quick 2> gpurun_out/c2n1.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['ms_per_step'],d['set_interp_data'])"
grep "cx_zone_create\|_setInterpDataChimera" gpurun_out/c2n1.err | tail -8 | cut -c1-250
