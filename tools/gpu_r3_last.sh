# round 2, session 3: last validation on one GPU (suite incl. the batched-hooks test), then smoke + default bench
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r03_g_bench_n1.json 2> gpurun_out/r03_g_bench_n1.err
grep -i "Traceback\|error" gpurun_out/r03_g_bench_n1.err | tail -3
python -c "
import json
d=json.loads(open('gpurun_out/r03_g_bench_n1.json').read().strip().splitlines()[-1])
print('c2a', d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['ms_per_step'],d['steps'],d['warmup']); s=d['set_interp_data']; print({k:s[k] for k in ('ms','search_kernel_ms','adt_build_ms')}); c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['e2e']['ms_per_step'],c['set_interp_data']['s'])"
