# round 2, session 3: 8-GPU bench line (C2 bodies + C4/C5 with transfer_host, batched hooks, queued searches)
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
CX_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r03_d_bench_n8.json 2> gpurun_out/r03_d_bench_n8.err
grep -i "error\|Traceback\|Mpi._setInterpData" gpurun_out/r03_d_bench_n8.err | head -8
python -c "
import json
d=json.loads(open('gpurun_out/r03_d_bench_n8.json').read().strip().splitlines()[-1])
print('c2n', d['value'],d['ms_per_step'],d['e2e'], d['set_interp_data']['s'], d['check']['ok']); c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['roofline']['frac'],c['e2e'],c['set_interp_data']['s'],c['check']['ok'],c['c5_100_iterations'])"
