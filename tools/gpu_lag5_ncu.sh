# order-5 solve kernel: one full capture (after the plain run exited 0)
cd $GRAFT_REPO_ROOT
timeout 300 python tests/prof_lagrange.py 5 128 2>&1 | tail -4 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lag_solve5r -c 1 -o gpurun_out/lag5r_v3 python tests/prof_lagrange.py 5 128 > gpurun_out/lag5_ncu2.log 2>&1
ls -la gpurun_out/
