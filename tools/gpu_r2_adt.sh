# round 2: ADT subtree pass over per-depth lists: parity + build times (C2a 16.6 M cells, C4 zones)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c1 or c2" 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --quick --sid-steps 2 --sid-warmup 1 > gpurun_out/r2_adt_c2a.json 2> gpurun_out/r2_adt_c2a.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_adt_c2a.json').read().strip().splitlines()[-1]); s=d['set_interp_data']; print('C2a adt_build_ms', s['adt_build_ms'], 'zone_create_s', s['zone_create_s'], 'sid ms', s['ms'])"
CX_PROFILE=1 CX_PROFILE_ZONE=1 timeout 300 python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/r2_adt_c4.json 2> gpurun_out/r2_adt_c4.err
grep "cx_zone_create" gpurun_out/r2_adt_c4.err | tail -1; grep "_setInterpDataChimera\]" gpurun_out/r2_adt_c4.err | tail -1
python -c "
import json
d=json.loads(open('gpurun_out/r2_adt_c4.json').read().strip().splitlines()[-1]); print('c4 sid', d['set_interp_data']['s'])"
