cd $GRAFT_REPO_ROOT
CX_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 5 --warmup 3 --quick > gpurun_out/bench_n2q.json 2> gpurun_out/bench_n2q.err
grep "^\[" gpurun_out/bench_n2q.err | head -20
