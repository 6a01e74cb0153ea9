# round 2: 2-GPU bench line with the e2e phase breakdown
cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
grep -i "error\|Traceback" gpurun_out/r2_bench_n2.err | head -5
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('c2n', d['value'],d['ms_per_step'],d['e2e']); c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['e2e'],c['set_interp_data']['s'])"
nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"
