set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --steps 2 --warmup 1 --no-cpu > gpurun_out/b18_$name.json 2> gpurun_out/b18_$name.err; }
run cfg5 cfg5 X=1
run cfg5_m3 cfg5 MPC_COOP_MINB=3
timeout 600 python -m pytest tests/test_gpu_boxqp.py -q -k "cfg5 or 124 or edge or inside" > gpurun_out/pytest18.log 2>&1; tail -3 gpurun_out/pytest18.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b18_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'), d.get('solved_only',{}).get('value'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
