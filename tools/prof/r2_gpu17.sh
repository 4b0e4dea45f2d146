set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rti.py tests/test_gpu_round2.py -q > gpurun_out/pytest17.log 2>&1; tail -5 gpurun_out/pytest17.log
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --steps 2 --warmup 1 --no-cpu > gpurun_out/b17_$name.json 2> gpurun_out/b17_$name.err; }
run cfg4 cfg4 X=1
run cfg4_f32 cfg4 X=1 --dtype f32
run obstacle obstacle X=1 --steps 1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b17_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'), d['e2e']['value'], d.get('summary'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
