# round 2, session 3: GPU suite (new preparer + transfer_host tests), then the 2-GPU bench line with transfer_host()
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for mode in host seq; do
CX_E2E=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3_bench_n2_$mode.json 2> gpurun_out/r3_bench_n2_$mode.err
grep -i "error\|Traceback" gpurun_out/r3_bench_n2_$mode.err | head -5
python -c "
import json
d=json.loads(open('gpurun_out/r3_bench_n2_$mode.json').read().strip().splitlines()[-1])
print('$mode c2n', d['value'],d['ms_per_step'],d['e2e']); c=d['c4_c5']; print('$mode c4', c['value'],c['ms_per_step'],c['e2e'],c['set_interp_data']['s'])"
done
