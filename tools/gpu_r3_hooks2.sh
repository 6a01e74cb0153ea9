# round 2, session 3: zone blocks from the stream-ordered pool -- parity suite, then the C4 tree-level call five times
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2 3; do
CX_PROFILE=1 timeout 600 python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/r3_hooks.json 2> gpurun_out/r3_hooks.err
grep -i "Traceback\|error" gpurun_out/r3_hooks.err | tail -3
grep _setInterpDataChimera gpurun_out/r3_hooks.err | tail -1
python -c "
import json
c=json.loads(open('gpurun_out/r3_hooks.json').read().strip().splitlines()[-1])
c=c.get('c4_c5',c)
print('   sid', c['set_interp_data']['s'], c['set_interp_data']['s_first_call'], c['check']['ok'])"
done
