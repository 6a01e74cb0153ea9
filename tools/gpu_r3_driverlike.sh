# what the driver runs at round end, on the final code: smoke, reference arm, default bench
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 1500 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r03_i_bench_ref_n1.json 2> gpurun_out/r03_i_bench_ref_n1.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r03_i_bench_n1.json 2> gpurun_out/r03_i_bench_n1.err; echo "bench rc=$?"
python -c "
import json
r=json.loads(open('gpurun_out/r03_i_bench_ref_n1.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/r03_i_bench_n1.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['e2e']['value'], r['config']==d['config'], r['metric']==d['metric'])
print('c2a', d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['ms_per_step'],d['gpu_launches'],d['clocks'])
c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['roofline']['frac'],c['e2e']['ms_per_step'],c['set_interp_data']['s'])
print('e2e ratio', d['e2e']['value']/r['e2e']['value'], 'device ratio', d['value']/r['value'])"
