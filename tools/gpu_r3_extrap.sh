# round 2, session 3: grouped extrapolation (k_extrap_zone + k_extrap_combine) against the warp-per-receptor form
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for m in group warp; do
CX_EXTRAP=$m CX_PROFILE=1 timeout 600 python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/r3_x_$m.json 2> gpurun_out/r3_x_$m.err
grep -i "_setInterpDataChimera\|Traceback\|error" gpurun_out/r3_x_$m.err | tail -3
python -c "
import json
c=json.loads(open('gpurun_out/r3_x_$m.json').read().strip().splitlines()[-1])
c=c.get('c4_c5',c)
print('$m c4', c['set_interp_data'], c['check']['ok'])"
done
