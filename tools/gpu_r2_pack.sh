# round 2: packed type-2 groups: parity of the transfer kernels, then A/B on C2a (N=1) and on the C4 step
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "transfer or pytree or smoke" 2>&1 | tail -4
for pk in 1 0; do
  CX_PLAN_PACK=$pk timeout 600 python bench.py --steps 20 --warmup 5 --quick --sid-steps 1 --sid-warmup 1 > gpurun_out/r2_pack_c2a_$pk.json 2> gpurun_out/r2_pack_c2a_$pk.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_pack_c2a_$pk.json").read().strip().splitlines()[-1])
    print("C2a PACK=$pk transfer ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"])
except Exception as e:
    print("C2a PACK=$pk failed", e); print(open("gpurun_out/r2_pack_c2a_$pk.err").read()[-2000:])
PY
  for st in 1 auto; do
    if [ $st = auto ]; then unset CX_PLAN_SORT; else export CX_PLAN_SORT=$st; fi
    CX_PLAN_PACK=$pk timeout 600 python bench.py --gpus 1 --workload c4 --steps 50 --warmup 5 --quick > gpurun_out/r2_pack_c4_${pk}_$st.json 2> gpurun_out/r2_pack_c4_${pk}_$st.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_pack_c4_${pk}_$st.json").read().strip().splitlines()[-1])
    print("C4 PACK=$pk SORT=$st us/step", 1e3*d["ms_per_step"], "frac", d["roofline"]["frac"], "c5", d["c5_100_iterations"]["ms_total"])
except Exception as e:
    print("C4 PACK=$pk SORT=$st failed", e); print(open("gpurun_out/r2_pack_c4_${pk}_$st.err").read()[-2000:])
PY
  done
  unset CX_PLAN_SORT
done
