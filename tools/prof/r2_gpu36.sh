cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=b36
timeout 900 python -m pytest tests/test_gpu_boxqp.py tests/test_gpu_round2.py -q -x -k "cfg5 or 124 or edge or inside or lti_4x" 2>&1 | tail -3
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --no-cpu > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; }
run cfg5 cfg5 X=1 --steps 2 --warmup 1
run cfg5_m3 cfg5 MPC_COOP_MINB=3 --steps 2 --warmup 1
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'), (d.get('solved_only') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
