# round 2: 2-GPU bench line (bounded by timeouts: a crashed rank must not leave the other one waiting)
cd $GRAFT_REPO_ROOT
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
grep -i "error\|Traceback" gpurun_out/r2_bench_n2.err | head -5
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('c2n', d['value'],d['ms_per_step'],d['e2e'], d['check']); c=d['c4_c5']; print('c4', c['value'],c['ms_per_step'],c['e2e'],c['set_interp_data']['s'], c['check'], c['c5_100_iterations'])"
timeout 300 python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | tail -3
