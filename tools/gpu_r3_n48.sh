# round 2, session 3: N=4 and N=8 bench lines of the final code (NCCL allgather of small objects on)
cd $GRAFT_REPO_ROOT
for n in 8; do
CX_PROFILE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r03_g_bench_n$n.json 2> gpurun_out/r03_g_bench_n$n.err
echo "N=$n rc=$?"
grep -i "error\|Traceback" gpurun_out/r03_g_bench_n$n.err | head -5
grep "Mpi._setInterpData" gpurun_out/r03_g_bench_n$n.err | tail -2
python -c "
import json
d=json.loads(open('gpurun_out/r03_g_bench_n$n.json').read().strip().splitlines()[-1])
print('c2n', d['value'],d['ms_per_step'],d['e2e']['ms_per_step'], d['set_interp_data']['s'], d['check']['ok']); c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['roofline']['frac'],c['e2e']['ms_per_step'],c['set_interp_data']['s'],c['set_interp_data'].get('calls_s_max_over_ranks'),c['check']['ok'],c['c5_100_iterations']['ms_total'])"
done
