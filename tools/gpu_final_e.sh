# final N=1 lines of the round (both arms), then launch list + full captures of the two dominant kernels
cd $GRAFT_REPO_ROOT
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -2 gpurun_out/bench_n1.err
wc -l gpurun_out/bench_n1.json gpurun_out/bench_ref_n1.json
python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/plain_f.json 2> gpurun_out/plain_f.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_f.csv python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_interp_o2 -c 1 -f -o gpurun_out/prof_search_f python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_transfer2s -s 3 -c 1 -f -o gpurun_out/prof_transfer_f python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ncu_f3.log 2>&1
python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/plain_c4f.json 2> gpurun_out/plain_c4f.err && \
ncu --set full --clock-control none -k regex:k_transfer_set -s 6 -c 1 -f -o gpurun_out/prof_c4_set_f python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/ncu_f4.log 2>&1
ls -la gpurun_out/*_f.ncu-rep
