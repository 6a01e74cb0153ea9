# ADT build with batched levels: parity (ADT replica, candidate lists, searches), then build times
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "adt or c1 or c2 or multizone or empty or golden or two_dim or extract" 2>&1 | tail -3
timeout 300 python tools/exp_zone_create.py 2>&1 | tail -7
timeout 600 python bench.py --steps 5 --warmup 3 --quick --sid-steps 2 --sid-warmup 1 > gpurun_out/adt_q.json 2> gpurun_out/adt_q.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/adt_q.json").read().strip().splitlines()[-1])
s=d["set_interp_data"]; print("C2a adt_build_ms", s["adt_build_ms"], "depth", s["adt_depth"], "zone_create_s", s["zone_create_s"], "sid ms", s["ms"])
PY
