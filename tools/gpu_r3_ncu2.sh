# round 2, session 3: C4 launch list + full capture of the grouped extrapolation kernel
cd $GRAFT_REPO_ROOT
C4="--gpus 1 --workload c4 --steps 5 --warmup 3 --quick"
python bench.py $C4 > gpurun_out/r03b_plain_c4.json 2> gpurun_out/r03b_plain_c4.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r03b_launches_c4.csv python bench.py $C4 > gpurun_out/r03b_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extrap_zone -s 2 -c 1 -f -o gpurun_out/r03b_extrap_zone_c4 python bench.py $C4 > gpurun_out/r03b_ncu6.log 2>&1
ls -la gpurun_out/r03b_*.ncu-rep
