# parity tests + A/B of transfer plan layouts (CX_PLAN_SORT: 1 donor-sorted, 2 tiled, unset auto)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
CX_PLAN_SORT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "transfer" 2>&1 | tail -2
CX_PLAN_SORT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "transfer" 2>&1 | tail -2
for st in 2 1; do
  CX_PLAN_SORT=$st python bench.py --steps 20 --warmup 5 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ab_sort$st.json 2> gpurun_out/ab_sort$st.err
  python -c "import json;d=json.load(open('gpurun_out/ab_sort$st.json'));print('sort',$st,d['ms_per_step'],d['value'],d['roofline']['frac'])"
done
python bench.py --steps 20 --warmup 5 --nvars 7 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ab_auto_7v.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/ab_auto_7v.json'));print('auto nvars 7',d['ms_per_step'],d['value'],d['roofline']['frac'])"
python bench.py --gpus 1 --workload c4 --steps 20 --warmup 5 --quick > gpurun_out/c4_n1_b.json 2> gpurun_out/c4_n1_b.err
python -c "import json;d=json.load(open('gpurun_out/c4_n1_b.json'));print('c4',d['ms_per_step'],d['value'])"
