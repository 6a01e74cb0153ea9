cd $GRAFT_REPO_ROOT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
grep -i "error\|Traceback" gpurun_out/bench_n4.err | head -5
wc -l gpurun_out/bench_n4.json
python -c "
import json
d=json.load(open('gpurun_out/bench_n4.json'))
print(d['value'],d['ms_per_step'],d['config']['transport'][:12], d['set_interp_data']['s'], d['e2e']['value']);c=d['c4_c5'];print(c['value'],c['ms_per_step'],c['c5_100_iterations'],c['set_interp_data']['s'])"
