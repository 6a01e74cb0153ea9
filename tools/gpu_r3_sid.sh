# round 2, session 3: async searches + single packing pass + batched ADT builds; D2H rate experiment
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/exp_d2h.py 2>&1 | tail -6
CX_PROFILE=1 timeout 600 python bench.py --gpus 1 --workload c4 --steps 20 --warmup 5 > gpurun_out/r3_c4_n1.json 2> gpurun_out/r3_c4_n1.err
grep -i "setInterpData\|Traceback\|error" gpurun_out/r3_c4_n1.err | tail -12
python -c "
import json
c=json.loads(open('gpurun_out/r3_c4_n1.json').read().strip().splitlines()[-1])
c=c.get('c4_c5',c)
print('c4', c['value'],c['ms_per_step'],c['e2e'],c['set_interp_data'])"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err
grep -i "Traceback\|error" gpurun_out/r3_bench_n1.err | tail -5
python -c "
import json
d=json.loads(open('gpurun_out/r3_bench_n1.json').read().strip().splitlines()[-1])
print('c2a', d['value'],d['ms_per_step'],d['e2e'],d['roofline']); print(d['set_interp_data']); c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['e2e'],c['set_interp_data'])"
