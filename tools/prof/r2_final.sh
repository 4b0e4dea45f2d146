set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/r02_bench_default_1gpu.json 2> gpurun_out/r02_bench_default_1gpu.err ) 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 300 python bench.py --workload cfg3 --horizon 20 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_cfg3_N20.json 2>/dev/null
timeout 300 python bench.py --workload cfg3 --dtype f32 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_cfg3_f32.json 2>/dev/null
timeout 300 python bench.py --workload obstacle --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_8f_obstacle_fused.json 2>/dev/null
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_nocpu.json 2> gpurun_out/r02_bench_nocpu.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_default_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('plan_only',{}).get('value'), d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'])
for k,v in d['secondary'].items(): print(k, v.get('value'), v.get('ms_per_step'), v['roofline'].get('traffic'), (v['roofline'].get('workspace_stream') or {}).get('frac'), v['roofline']['fp_pipe']['frac'], (v.get('solved_only') or {}).get('value'))
for f in ('r02_bench_cfg3_N20','r02_bench_cfg3_f32','r02_bench_8f_obstacle_fused','r02_bench_reference_arm'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('ms_per_step'))
    except Exception as e: print(f,'ERR',e)
P
