# round 2: box layout of the fringe plans: parity of the transfer kernels, then A/B on the C4 step (N=1, 8 zones, 7 vars)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "transfer" 2>&1 | tail -4
for st in 1 auto 5; do
  if [ $st = auto ]; then unset CX_PLAN_SORT; else export CX_PLAN_SORT=$st; fi
  timeout 600 python bench.py --gpus 1 --workload c4 --steps 50 --warmup 5 --quick > gpurun_out/c4ab_$st.json 2> gpurun_out/c4ab_$st.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c4ab_$st.json").read().strip().splitlines()[-1])
    print("CX_PLAN_SORT=$st", "us/step", 1e3*d["ms_per_step"], "frac", d["roofline"]["frac"], "c5", d["c5_100_iterations"]["ms_total"], "e2e ms", d["e2e"]["ms_per_step"], "sid s", d["set_interp_data"]["s"])
except Exception as e:
    print("CX_PLAN_SORT=$st failed", e); print(open("gpurun_out/c4ab_$st.err").read()[-2000:])
PY
done
