# whole GPU suite on one GPU, then the quick C2a bench legs
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 5 --quick --sid-steps 2 --sid-warmup 1 > gpurun_out/quick_n1.json 2> gpurun_out/quick_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/quick_n1.json").read().strip().splitlines()[-1])
s=d["set_interp_data"]; print(d["ms_per_step"], d["roofline"]["frac"], "adt_build_ms", s["adt_build_ms"], "sid ms", s["ms"])
PY
