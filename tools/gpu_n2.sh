# 2-GPU: full GPU suite (incl. NCCL/P2P parity tests) + bench line (both arms)
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "NOT YET" | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep -i "error\|Traceback" gpurun_out/bench_n2.err | head -5
wc -l gpurun_out/bench_n2.json
python -c "
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print(d['value'],d['ms_per_step'],d['config']['transport'][:12], d['set_interp_data']['s'], d['set_interp_data']['s_first_call']);c=d['c4_c5'];print(c['value'],c['ms_per_step'],c['c5_100_iterations'],c['set_interp_data']['s'])"
