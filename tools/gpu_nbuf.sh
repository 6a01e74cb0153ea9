# staged type-2 kernel: 2 / 3 / 4 variables staged per CTA (variant builds), C2a with 5 and 7 variables + C4 step
cd $GRAFT_REPO_ROOT
for v in "" _nb3 _nb4; do
  export CX_LIBPATH=$GRAFT_REPO_ROOT/cassiopee_b200/csrc/libcassiopee_b200$v.so
  for nv in 5 7; do
    timeout 600 python bench.py --steps 30 --warmup 5 --quick --nvars $nv --sid-steps 1 --sid-warmup 0 > gpurun_out/nbuf$v_$nv.json 2> gpurun_out/nbuf$v_$nv.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/nbuf$v_$nv.json").read().strip().splitlines()[-1])
print("lib$v nvars=$nv", d["ms_per_step"], d["value"], d["roofline"]["frac"])
PY
  done
  timeout 600 python bench.py --gpus 1 --workload c4 --steps 50 --warmup 10 --quick > gpurun_out/nbufc4$v.json 2> gpurun_out/nbufc4$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/nbufc4$v.json").read().strip().splitlines()[-1])
print("lib$v C4", d["ms_per_step"], d["value"])
PY
done
