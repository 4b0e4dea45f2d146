cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=b32
timeout 600 python -m pytest tests/test_gpu_rti.py tests/test_gpu_boxqp.py -q -x 2>&1 | tail -2
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --no-cpu > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; }
run cfg4_pf1 cfg4 MPC_QP_PREFETCH=1 --steps 2 --warmup 1
run cfg4_pf2 cfg4 MPC_QP_PREFETCH=2 --steps 2 --warmup 1
run cfg4_pf3 cfg4 MPC_QP_PREFETCH=3 --steps 2 --warmup 1
run cfg3_nostage_pf1 cfg3 "MPC_QP_STAGED=0" --steps 5 --warmup 3
run cfg3_nostage_pf2 cfg3 "MPC_QP_STAGED=0 MPC_QP_PREFETCH=2" --steps 5 --warmup 3
run obstacle_pf2 obstacle MPC_QP_PREFETCH=2 --steps 2 --warmup 1
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
