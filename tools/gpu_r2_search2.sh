cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest "tests/test_gpu_fullsize.py::test_c1_fullsize_bit_exact_vs_reference" -x -q -m gpu 2>&1 | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_interp_o2 -c 1 -f -o gpurun_out/r2_interp_o2 \
   python bench.py --steps 2 --warmup 3 --quick --sid-steps 1 --sid-warmup 0 > gpurun_out/r2_ncu_interp.log 2>&1
echo "ncu rc=$?"
