# 2-GPU validation: whole GPU suite (incl. the distributed tests), then a short N=2 bench line
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("N=2", d["ms_per_step"], d["value"], d["roofline"]["frac"], json.dumps(d.get("c4_c5"))[:600])
PY
