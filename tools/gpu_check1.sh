# 1-GPU regression: parity tests, C4 single-GPU step, tiled-layout A/B, default bench line
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --gpus 1 --workload c4 --steps 20 --warmup 5 --quick > gpurun_out/c4_n1_c.json 2> gpurun_out/c4_n1_c.err
python -c "import json;d=json.load(open('gpurun_out/c4_n1_c.json'));print('c4',d['ms_per_step'],d['value'],d['c5_100_iterations'],d['set_interp_data']['s'])"
CX_PLAN_SORT=2 python bench.py --steps 20 --warmup 5 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ab_sort2b.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/ab_sort2b.json'));print('tiled pitch34',d['ms_per_step'],d['value'])"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -2 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
