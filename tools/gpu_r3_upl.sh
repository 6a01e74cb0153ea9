# round 2, session 3: staging threads of pageable uploads on the C4 tree-level call
cd $GRAFT_REPO_ROOT
for t in 8 16 4 12; do
CX_UPLOAD_THREADS=$t CX_PROFILE=1 timeout 600 python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/r3_upl.json 2> gpurun_out/r3_upl.err
echo "threads=$t $(grep _setInterpDataChimera gpurun_out/r3_upl.err | tail -1)"
python -c "
import json
c=json.loads(open('gpurun_out/r3_upl.json').read().strip().splitlines()[-1])
c=c.get('c4_c5',c)
print('   sid', c['set_interp_data']['s'], c['set_interp_data'].get('calls_s_max_over_ranks'))"
done
nproc
