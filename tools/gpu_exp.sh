cd $GRAFT_REPO_ROOT
for tb in 256 512 256 512; do echo TB=$tb; CX_T2S_TB=$tb python tools/exp_store_modes.py 2>&1 | tail -3 | head -1; done
