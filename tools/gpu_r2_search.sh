# round 2: search kernel without local memory: parity, then A/B of the CTA size on C2a (17.0 M receptors)
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for tb in 128 64; do
  CX_SEARCH_TB=$tb timeout 600 python bench.py --steps 5 --warmup 3 --quick --sid-steps 3 --sid-warmup 1 > gpurun_out/r2_search_tb$tb.json 2> gpurun_out/r2_search_tb$tb.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_search_tb$tb.json").read().strip().splitlines()[-1]); s=d["set_interp_data"]
    print("TB=$tb sid ms", s["ms"], "search kernel ms", s["search_kernel_ms"], "nodes/query", s["nodes_visited_per_query"], "transfer ms", d["ms_per_step"])
except Exception as e:
    print("TB=$tb failed", e); print(open("gpurun_out/r2_search_tb$tb.err").read()[-2000:])
PY
done
timeout 1700 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
