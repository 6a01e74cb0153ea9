# round 2, session 3: receptor points taken from the resident hook (tree interpolated from itself at the nodes)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | tail -3
CX_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03_k_bench_n2.json 2> gpurun_out/r03_k_bench_n2.err
grep -i "error\|Traceback" gpurun_out/r03_k_bench_n2.err | head -5
grep "Mpi._setInterpData\|_setInterpDataChimera" gpurun_out/r03_k_bench_n2.err | sed -n 3,6p
python -c "
import json
d=json.loads(open('gpurun_out/r03_k_bench_n2.json').read().strip().splitlines()[-1])
print('c2n', d['value'],d['ms_per_step'],d['e2e']['ms_per_step'], d['set_interp_data']['s'], d['check']['ok']); c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['e2e']['ms_per_step'],c['set_interp_data']['s'],c['check']['ok'])"
