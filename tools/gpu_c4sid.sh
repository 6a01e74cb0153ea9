# phase profile of the tree-level setInterpData on the C4 blocks of one GPU (CX_PROFILE / CX_PROFILE_ZONE)
cd $GRAFT_REPO_ROOT
CX_PROFILE=1 CX_PROFILE_ZONE=1 timeout 600 python bench.py --gpus 1 --workload c4 --steps 3 --warmup 2 --quick > gpurun_out/c4sid.json 2> gpurun_out/c4sid.err
grep "cx_zone_create\|_setInterpDataChimera\|Mpi._setInterpData" gpurun_out/c4sid.err | tail -14
