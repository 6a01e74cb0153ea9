# round 2, session 3: searches on side streams; zone-create profile; N=2 e2e with explicit host threads
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
CX_PROFILE=1 CX_PROFILE_ZONE=1 timeout 600 python bench.py --gpus 1 --workload c4 --steps 20 --warmup 5 > gpurun_out/r3_c4_n1.json 2> gpurun_out/r3_c4_n1.err
grep -i "cx_zone_create\|setInterpData\|Traceback\|error" gpurun_out/r3_c4_n1.err | tail -14
python -c "
import json
c=json.loads(open('gpurun_out/r3_c4_n1.json').read().strip().splitlines()[-1])
c=c.get('c4_c5',c)
print('c4', c['value'],c['ms_per_step'],c['e2e'],c['set_interp_data'])"
