# round 2: lean plan-set kernel A/B on the C4 step (N=1, 8 zones, 7 vars) + the sparse-download test
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sparse_download or planset or plan_set or box_layout or packed" 2>&1 | tail -3
for cfg in "1 0" "1 6" "1 8" "auto 0"; do
  set -- $cfg
  if [ $1 = auto ]; then unset CX_PLAN_SORT; else export CX_PLAN_SORT=$1; fi
  CX_SET_LEAN=$2 timeout 300 python bench.py --gpus 1 --workload c4 --steps 50 --warmup 5 --quick > gpurun_out/r2_lean_$1_$2.json 2> gpurun_out/r2_lean_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_lean_$1_$2.json").read().strip().splitlines()[-1])
    print("C4 SORT=$1 LEAN=$2 us/step", 1e3*d["ms_per_step"], "frac", d["roofline"]["frac"], "c5", d["c5_100_iterations"]["ms_total"], "check", d["check"]["ok"], d["check"]["max_abs_err"])
except Exception as e:
    print("C4 SORT=$1 LEAN=$2 failed", e); print(open("gpurun_out/r2_lean_$1_$2.err").read()[-1500:])
PY
done
