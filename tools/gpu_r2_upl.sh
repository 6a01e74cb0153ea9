# round 2: threaded pageable uploads A/B on the C4 tree-level setInterpData (N=1)
cd $GRAFT_REPO_ROOT
for th in 1 4 8; do
  CX_UPLOAD_THREADS=$th CX_PROFILE=1 CX_PROFILE_ZONE=1 timeout 300 python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/r2_upl_$th.json 2> gpurun_out/r2_upl_$th.err
  echo "threads=$th"; grep "cx_zone_create" gpurun_out/r2_upl_$th.err | tail -2; grep "_setInterpDataChimera\]" gpurun_out/r2_upl_$th.err | tail -1
  python -c "
import json
d=json.loads(open('gpurun_out/r2_upl_$th.json').read().strip().splitlines()[-1]); print('sid', d['set_interp_data']['s'], d['set_interp_data']['s_first_call'])"
done
