set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T=${TAG:-b26}
timeout 900 python -m pytest tests/test_gpu_boxqp.py tests/test_gpu_round2.py -q -x > gpurun_out/pytest_$T.log 2>&1; tail -5 gpurun_out/pytest_$T.log
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --no-cpu > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; }
run cfg3 cfg3 X=1 --steps 5 --warmup 3
run cfg3_nostage cfg3 MPC_QP_STAGED=0 --steps 5 --warmup 3
run cfg3_m5 cfg3 MPC_QP_MINB=5 --steps 5 --warmup 3
run cfg3_m6 cfg3 MPC_QP_MINB=6 --steps 5 --warmup 3
run cfg3_f32 cfg3 X=1 --steps 5 --warmup 3 --dtype f32
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
