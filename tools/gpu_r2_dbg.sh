# round 2: compact-transfer GPU parity; the N=2 C4 plan shapes on ONE process (all 16 zones local), plain and under compute-sanitizer
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_compact_transfers.py tests/test_gpu_parity.py -x -q -m gpu -k "compact or transfer" 2>&1 | tail -4
timeout 300 python bench.py --gpus 2 --workload c4 --steps 10 --warmup 3 --quick > gpurun_out/r2_dbg_c4.json 2> gpurun_out/r2_dbg_c4.err; echo "c4 16 zones plain rc=$?"
grep -i "error\|illegal" gpurun_out/r2_dbg_c4.err | head -5
python -c "
import json
d=json.loads(open('gpurun_out/r2_dbg_c4.json').read().strip().splitlines()[-1]); print('us/step', 1e3*d['ms_per_step'], d['config']['plans'], d['config']['receptors'])"
CX_C4_N=40 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --gpus 2 --workload c4 --steps 2 --warmup 1 --quick > gpurun_out/r2_dbg_san.out 2>&1; echo "sanitizer rc=$?"
grep -B2 -A12 "Invalid\|ERROR SUMMARY" gpurun_out/r2_dbg_san.out | head -60
