cd $GRAFT_REPO_ROOT
CX_TRANSFER_STAGING=0 python tests/prof_lagrange.py 5 1 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -2 gpurun_out/bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']);print(json.dumps(d['other_configs'],indent=1))"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
grep -v "NOT YET" gpurun_out/bench_ref_n1.err | tail -2
python -c "import json;d=json.load(open('gpurun_out/bench_ref_n1.json'));print(d['value'],d['set_interp_data']);print(json.dumps(d['other_configs'],indent=1))"
