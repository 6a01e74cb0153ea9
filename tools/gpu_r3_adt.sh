# round 2, session 3: compact children array in the ADT build -- parity suite, then C2a / C4 numbers
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3_adt_n1.json 2> gpurun_out/r3_adt_n1.err
grep -i "Traceback\|error" gpurun_out/r3_adt_n1.err | tail -5
python -c "
import json
d=json.loads(open('gpurun_out/r3_adt_n1.json').read().strip().splitlines()[-1])
print('c2a', d['value'],d['ms_per_step'],d['e2e']['ms_per_step']); s=d['set_interp_data']; print({k:s[k] for k in ('ms','search_kernel_ms','adt_build_ms','zone_create_s')}, s['e2e']['calls'][-1]); c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['e2e']['ms_per_step'],c['set_interp_data']['s'])"
