# round 2: new GPU tests + the N=1 bench line (both arms)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_compact_transfers.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench n1 rc=$?"
grep -i "error\|Traceback" gpurun_out/r2_bench_n1.err | head -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('c2a', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
s=d['set_interp_data']; print('sid ms', s['ms'], 'kernel', s['search_kernel_ms'], 'e2e', s['e2e']['value'], [c['total_s'] for c in s['e2e']['calls']])
for k,v in d['other_configs'].items(): print(k, json.dumps(v)[:300])
c=d['c4_c5']; print('c4', c.get('value'), c.get('ms_per_step'), c.get('e2e'), c.get('set_interp_data'), c.get('check'), c.get('c5_100_iterations'))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref_n1.json 2> gpurun_out/r2_bench_ref_n1.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_ref_n1.json').read().strip().splitlines()[-1])
print('ref', d['value'], d['ms_per_step'], d['steps'], d['warmup'], d['set_interp_data'], d['config'])
PY
