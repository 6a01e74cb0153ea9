# round 2, session 3 evidence: launch lists of the bench commands + full captures of the dominant kernels (each after the
# plain run of the same command has exited 0)
cd $GRAFT_REPO_ROOT
Q="--steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0"
python bench.py $Q > gpurun_out/r03_plain_c2a.json 2> gpurun_out/r03_plain_c2a.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r03_launches_c2a.csv python bench.py $Q > gpurun_out/r03_ncu1.log 2>&1
tail -1 gpurun_out/r03_ncu1.log | cut -c1-160
ncu --set full --clock-control none --import-source on -k regex:k_transfer2s -s 3 -c 1 -f -o gpurun_out/r03_transfer2s python bench.py $Q > gpurun_out/r03_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pack_scatter -c 1 -f -o gpurun_out/r03_pack_scatter python bench.py $Q > gpurun_out/r03_ncu2.log 2>&1
C4="--gpus 1 --workload c4 --steps 5 --warmup 3 --quick"
python bench.py $C4 > gpurun_out/r03_plain_c4.json 2> gpurun_out/r03_plain_c4.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r03_launches_c4.csv python bench.py $C4 > gpurun_out/r03_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_transfer_set -s 20 -c 1 -f -o gpurun_out/r03_transfer_set_c4 python bench.py $C4 > gpurun_out/r03_ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extrap -s 2 -c 1 -f -o gpurun_out/r03_extrap_c4 python bench.py $C4 > gpurun_out/r03_ncu6.log 2>&1
ls -la gpurun_out/r03_*.ncu-rep
