# round 2: memory floor of the C4 step's footprint (tools/exp/slab_floor.cu) + ncu --set full of the step kernel, sorted vs box layout
cd $GRAFT_REPO_ROOT
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/slab_floor tools/exp/slab_floor.cu 2>/dev/null && /tmp/slab_floor
for st in 1 5; do
  CX_PLAN_SORT=$st timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_transfer_set -s 20 -c 1 -f -o gpurun_out/r2_c4_sort$st \
     python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/r2_ncu_c4_sort$st.log 2>&1
  echo "ncu sort$st rc=$?"
done
