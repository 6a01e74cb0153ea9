cd $GRAFT_REPO_ROOT
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -2 gpurun_out/bench_n1.err
wc -l gpurun_out/bench_n1.json gpurun_out/bench_ref_n1.json
python -c "
import json
d=json.load(open('gpurun_out/bench_n1.json')); r=json.load(open('gpurun_out/bench_ref_n1.json'))
print('ours', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'clocks', d['clocks'])
print('sid', d['set_interp_data']['value'], d['set_interp_data']['ms'], d['set_interp_data']['e2e']['value'], d['set_interp_data']['adt_build_ms'], d['set_interp_data']['zone_create_s'])
print('cpu', d['cpu_baseline'])
print('ref', r['value'], r['set_interp_data']['value'])
for k,v in d['other_configs'].items(): print(k, json.dumps(v)[:300])
"
