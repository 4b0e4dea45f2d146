cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=b37
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --no-cpu > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; }
run cfg5_pf0 cfg5 MPC_COOP_PREFETCH=0 --steps 2 --warmup 1 --batch 524288
run cfg5_pf1 cfg5 MPC_COOP_PREFETCH=1 --steps 2 --warmup 1 --batch 524288
run cfg5_pf2 cfg5 MPC_COOP_PREFETCH=2 --steps 2 --warmup 1 --batch 524288
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'), (d.get('solved_only') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
