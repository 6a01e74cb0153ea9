# round 2, session 3: sparse donor upload in the host-pointer pipeline (C2a e2e), A/B + the new GPU test
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sparse_upload or transfer_matches or synthetic or batch_keeps" 2>&1 | tail -4
for m in 1 0; do
CX_SPARSE_H2D=$m timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3_e2e_$m.json 2> gpurun_out/r3_e2e_$m.err
grep -i "Traceback\|error" gpurun_out/r3_e2e_$m.err | tail -5
python -c "
import json
d=json.loads(open('gpurun_out/r3_e2e_$m.json').read().strip().splitlines()[-1])
print('sparse=$m c2a', d['value'],d['ms_per_step'],d['e2e'])"
done
