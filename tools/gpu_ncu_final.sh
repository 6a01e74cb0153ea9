# launch list of the default bench command (quick legs) + full captures of the two dominant kernels
cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/plain_d.json 2> gpurun_out/plain_d.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_d.csv python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ncu_d1.log 2>&1
tail -1 gpurun_out/ncu_d1.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:k_interp_o2 -c 1 -f -o gpurun_out/prof_search_d python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ncu_d2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_transfer2s -s 3 -c 1 -f -o gpurun_out/prof_transfer_d python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ncu_d3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
