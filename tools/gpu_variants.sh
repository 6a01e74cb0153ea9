# the transfer tests under every plan layout / kernel switch
cd $GRAFT_REPO_ROOT
for env in "CX_TRANSFER_STAGING=0" "CX_PLAN_SORT=0" "CX_PLAN_SORT=2" "CX_PLAN_SORT=3" "CX_PLAN_SORT=4" "CX_T2S_TB=128"; do
  echo "== $env"
  env $env timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "transfer or synthetic or lagrange_coincident or fullsize_adt or c4_cluster or two_dimensional or extract" 2>&1 | grep -v "NOT YET" | tail -2
done
