# 2-GPU: NCCL/P2P parity tests + multi-body bench line with both transports
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | grep -v "NOT YET" | tail -15
for tr in p2p nccl; do
CX_PROFILE=1 CX_TRANSPORT=$tr timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_$tr.json 2> gpurun_out/bench_n2_$tr.err
grep -i "transport\|error\|Traceback" gpurun_out/bench_n2_$tr.err | head -5
python -c "
import json
s=open('gpurun_out/bench_n2_$tr.json').read()
d=json.loads([l for l in s.splitlines() if l.startswith('{')][-1])
print('$tr', d['value'],d['ms_per_step'],d['config']['transport'][:12]);c=d['c4_c5'];print(c['value'],c['ms_per_step'],c['c5_100_iterations'])"
done
