# last check of the final tree: smoke and the default bench line
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r03_m_bench_n1.json 2> gpurun_out/r03_m_bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r03_m_bench_n1.json').read().strip().splitlines()[-1])
print('c2a', d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['ms_per_step']); c=d['c4_c5']; print('c4', c['ms_per_step'],c['e2e']['ms_per_step'],c['set_interp_data']['s'],c['check']['ok'])"
