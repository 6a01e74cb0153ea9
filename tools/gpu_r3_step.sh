# round 2, session 3: smoke(), then the C4 step time (k_transfer_set) after the plan-load change
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for i in 1 2; do
timeout 600 python bench.py --gpus 1 --workload c4 --steps 50 --warmup 10 --quick > gpurun_out/r3_step_$i.json 2> gpurun_out/r3_step_$i.err
python -c "
import json
c=json.loads(open('gpurun_out/r3_step_$i.json').read().strip().splitlines()[-1])
c=c.get('c4_c5',c)
print('c4 step', c['ms_per_step'], c['c5_100_iterations']['ms_per_step'], c['check']['ok'])"
done
