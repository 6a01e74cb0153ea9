# round 2, session 3: bounding-box upload in the host-pointer pipeline (C2a e2e), A/B + its GPU test
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sparse_upload or transfer_matches or synthetic" 2>&1 | tail -3
for m in 1 0 1 0; do
CX_BOX_H2D=$m timeout 900 python bench.py --steps 20 --warmup 5 --nvars 5 > gpurun_out/r3_box_$m.json 2> gpurun_out/r3_box_$m.err
grep -i "Traceback\|error" gpurun_out/r3_box_$m.err | tail -3
python -c "
import json
d=json.loads(open('gpurun_out/r3_box_$m.json').read().strip().splitlines()[-1])
print('box=$m c2a', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e'].get('donor_points_uploaded'))"
done
