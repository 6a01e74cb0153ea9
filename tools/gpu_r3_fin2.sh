# after the signature change of Xmpi._setInterpData: the callers on the GPU side still run (2 GPUs)
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c4" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 5 --warmup 3 --quick > gpurun_out/r03_l_bench_n2.json 2> gpurun_out/r03_l_bench_n2.err; echo "rc=$?"
grep -i "error\|Traceback" gpurun_out/r03_l_bench_n2.err | head -5
python -c "
import json
d=json.loads(open('gpurun_out/r03_l_bench_n2.json').read().strip().splitlines()[-1])
print('c2n', d['value'],d['ms_per_step'], d['set_interp_data']['s'], d['check']['ok'])"
