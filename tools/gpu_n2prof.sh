# 2-GPU: distributed parity tests (device zone replication), then the C4 setInterpData phase profile for both replication paths
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | tail -4
for mode in device host; do
  echo "== CX_ZONE_REPL=$mode"
  CX_ZONE_REPL=$mode CX_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/n2prof_$mode.json 2> gpurun_out/n2prof_$mode.err
  grep "Mpi._setInterpData\|_setInterpDataChimera\|Error\|error" gpurun_out/n2prof_$mode.err | tail -6
  python - <<PY
import json
d=json.loads(open("gpurun_out/n2prof_$mode.json").read().strip().splitlines()[-1])
print(d["set_interp_data"]["s"], d["set_interp_data"]["s_first_call"], d["ms_per_step"])
PY
done
