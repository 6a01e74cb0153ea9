# round 2, session 3, final evidence on one GPU: parity suite, both bench arms, launch lists of the final code
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r03_d_bench_n1.json 2> gpurun_out/r03_d_bench_n1.err
grep -i "Traceback\|error" gpurun_out/r03_d_bench_n1.err | tail -3
python -c "
import json
d=json.loads(open('gpurun_out/r03_d_bench_n1.json').read().strip().splitlines()[-1])
print('c2a', d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['ms_per_step'],d['cpu_baseline']['value']); s=d['set_interp_data']; print({k:s[k] for k in ('ms','search_kernel_ms','adt_build_ms')}, s['e2e']['calls'][-1]); c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['roofline'],c['e2e']['ms_per_step'],c['set_interp_data']['s'], c['set_interp_data'].get('calls_s_max_over_ranks'))"
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03_d_bench_ref_n1.json 2> gpurun_out/r03_d_bench_ref_n1.err
tail -c 600 gpurun_out/r03_d_bench_ref_n1.json
Q="--steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0"
python bench.py $Q > gpurun_out/r03_d_plain_c2a.json 2> gpurun_out/r03_d_plain_c2a.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r03_d_launches_c2a.csv python bench.py $Q > gpurun_out/r03_d_ncu1.log 2>&1
C4="--gpus 1 --workload c4 --steps 5 --warmup 3 --quick"
python bench.py $C4 > gpurun_out/r03_d_plain_c4.json 2> gpurun_out/r03_d_plain_c4.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r03_d_launches_c4.csv python bench.py $C4 > gpurun_out/r03_d_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extrap_zone -s 2 -c 1 -f -o gpurun_out/r03_d_extrap_zone_c4 python bench.py $C4 > gpurun_out/r03_d_ncu6.log 2>&1
ls -la gpurun_out/r03_d_*
