set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_rti.py -q > gpurun_out/pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest11.log
tail -12 gpurun_out/pytest11.log
timeout 900 python bench.py > gpurun_out/b11_default.json 2> gpurun_out/b11_default.err; echo "bench rc=$?"; tail -3 gpurun_out/b11_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b11_reference.json 2> gpurun_out/b11_reference.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/b11_default.json').read().strip().splitlines()[-1])
print('headline', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_host_link'], d['e2e']['host_link'], 'plan_only', d['e2e']['plan_only']['value'])
print('guard', d['roofline']['guard_divergence'])
print('cpu', d.get('cpu_baseline'))
for k,v in d['secondary'].items(): print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v.get('solved_only'))
P
