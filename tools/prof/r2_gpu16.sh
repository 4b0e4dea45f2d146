set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q -k "ordered or session23 or cfg5" > gpurun_out/pytest16.log 2>&1; tail -5 gpurun_out/pytest16.log
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --steps 5 --warmup 3 --no-cpu > gpurun_out/b16_$name.json 2> gpurun_out/b16_$name.err; }
run cfg3 cfg3 X=1
run cfg3_N20 cfg3 X=1 --horizon 20
run cfg3_f32 cfg3 X=1 --dtype f32
run cfg3_m3 cfg3 MPC_QP_MINB=3
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b16_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'), d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
