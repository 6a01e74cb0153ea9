cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:k_transfer_set -s 20 -c 2 -f -o gpurun_out/prof_c4_set python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/ncu_c4set.log 2>&1
tail -2 gpurun_out/ncu_c4set.log
