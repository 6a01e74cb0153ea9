# round 2, session 3: plan data with the L2 evict_last policy (CX_PLAN_KEEP = cap in MB; 0 = plain loads), C4 step
cd $GRAFT_REPO_ROOT
for k in 0 48 0 48; do
CX_PLAN_KEEP=$k timeout 600 python bench.py --gpus 1 --workload c4 --steps 50 --warmup 10 --quick > gpurun_out/r3_keep_$k.json 2> gpurun_out/r3_keep_$k.err
grep -i "Traceback\|error" gpurun_out/r3_keep_$k.err | tail -3
python -c "
import json
c=json.loads(open('gpurun_out/r3_keep_$k.json').read().strip().splitlines()[-1])
c=c.get('c4_c5',c)
print('keep=$k c4', c['value'],c['ms_per_step'],c.get('c5_100_iterations'),c['check']['ok'],c['check']['max_abs_err'])"
done
