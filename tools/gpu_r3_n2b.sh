# round 2, session 3: world-2 tests and the 2-GPU bench line with both allgather paths
cd $GRAFT_REPO_ROOT
CX_ALLGATHER=nccl timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | tail -3
for ag in gloo nccl; do
CX_ALLGATHER=$ag CX_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r03_d_bench_n2_$ag.json 2> gpurun_out/r03_d_bench_n2_$ag.err
grep -i "error\|Traceback" gpurun_out/r03_d_bench_n2_$ag.err | head -5
grep "Mpi._setInterpData" gpurun_out/r03_d_bench_n2_$ag.err | tail -2
python -c "
import json
d=json.loads(open('gpurun_out/r03_d_bench_n2_$ag.json').read().strip().splitlines()[-1])
print('$ag c2n', d['value'],d['ms_per_step'],d['e2e']['ms_per_step'], d['set_interp_data']['s'], d['check']['ok']); c=d['c4_c5']; print('$ag c4', c['value'],c['ms_per_step'],c['e2e']['ms_per_step'],c['set_interp_data']['s'],c['set_interp_data'].get('calls_s_max_over_ranks'),c['check']['ok'])"
done
