cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=b35
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --no-cpu > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; }
run cfg3_4cta cfg3 X=1 --steps 5 --warmup 3
run cfg3_3cta cfg3 MPC_QP_PAD_SMEM=30000 --steps 5 --warmup 3
run cfg3_2cta cfg3 MPC_QP_PAD_SMEM=65000 --steps 5 --warmup 3
run cfg3_1cta cfg3 MPC_QP_PAD_SMEM=150000 --steps 5 --warmup 3
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
