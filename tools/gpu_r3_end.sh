# final state: whole GPU suite, then the default bench line
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r03_j_bench_n1.json 2> gpurun_out/r03_j_bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r03_j_bench_n1.json').read().strip().splitlines()[-1])
print('c2a', d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['h2d_bytes_per_step'])
c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['roofline']['frac'],c['e2e']['ms_per_step'],c['set_interp_data']['s'])"
