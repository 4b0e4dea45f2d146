set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/r02_bench_default_1gpu.err ) 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_nocpu.json 2> gpurun_out/r02_bench_nocpu.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_default_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])
for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('ms_per_step'))
P
