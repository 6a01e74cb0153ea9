# round 2, session 3: residency / group size of the grouped extrapolation kernel on the C4 tree-level setInterpData
cd $GRAFT_REPO_ROOT
for cfg in "2 16" "4 16" "6 16" "4 8" "6 8" "6 4"; do
set -- $cfg
CX_XMINB=$1 CX_XG=$2 CX_PROFILE=1 timeout 600 python bench.py --gpus 1 --workload c4 --steps 5 --warmup 3 --quick > gpurun_out/r3_xg_$1_$2.json 2> gpurun_out/r3_xg_$1_$2.err
grep -i "Traceback\|error" gpurun_out/r3_xg_$1_$2.err | tail -3
echo "minb=$1 xg=$2 $(grep _setInterpDataChimera gpurun_out/r3_xg_$1_$2.err | tail -1)"
python -c "
import json
c=json.loads(open('gpurun_out/r3_xg_$1_$2.json').read().strip().splitlines()[-1])
c=c.get('c4_c5',c)
print('   sid', c['set_interp_data']['s'], c['check']['ok'])"
done
