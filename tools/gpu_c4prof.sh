set -x
cd $GRAFT_REPO_ROOT
python bench.py --gpus 1 --workload c4 --steps 20 --warmup 5 --quick > gpurun_out/c4_n1.json 2> gpurun_out/c4_n1.err
tail -3 gpurun_out/c4_n1.err
cat gpurun_out/c4_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c4.csv python bench.py --gpus 1 --workload c4 --steps 3 --warmup 3 --quick > gpurun_out/ncu_c4.log 2>&1
tail -3 gpurun_out/ncu_c4.log
