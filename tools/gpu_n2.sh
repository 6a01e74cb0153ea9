# 2-GPU: NCCL parity tests + multi-body bench line
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -5 gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json
