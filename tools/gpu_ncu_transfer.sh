# ncu --set full of the transfer kernel variants (one launch each)
cd $GRAFT_REPO_ROOT
for st in 2 1; do
CX_PLAN_SORT=$st ncu --set full --clock-control none --import-source on -k regex:k_transfer2 -s 3 -c 1 -f -o gpurun_out/prof_transfer_sort$st python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/ncu_sort$st.log 2>&1
tail -2 gpurun_out/ncu_sort$st.log
done
