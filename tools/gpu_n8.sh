# 8-GPU: multi-body bench line + C4/C5 (64 zones)
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -5 gpurun_out/bench_n8.err
cat gpurun_out/bench_n8.json
