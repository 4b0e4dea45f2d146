cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=b33
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --no-cpu > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; }
run cfg4_b37888 cfg4 X=1 --steps 1 --warmup 1 --batch 37888
run cfg4_b37888_1cta cfg4 MPC_RTI_PAD_SMEM=120000 --steps 1 --warmup 1 --batch 37888
run cfg4_b18944_1cta cfg4 MPC_RTI_PAD_SMEM=120000 --steps 1 --warmup 1 --batch 18944
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
