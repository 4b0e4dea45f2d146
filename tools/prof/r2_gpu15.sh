set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --steps 2 --warmup 2 --no-cpu > gpurun_out/b15_$name.json 2> gpurun_out/b15_$name.err; }
run cfg3_t128 cfg3 X=1
run cfg3_t64 cfg3 MPC_QP_THREADS=64
run cfg3_t32 cfg3 MPC_QP_THREADS=32
run cfg4_t128 cfg4 X=1
run cfg4_t64 cfg4 MPC_RTI_THREADS=64
run cfg4_t32 cfg4 MPC_RTI_THREADS=32
run obst_t32 obstacle MPC_RTI_THREADS=32 --steps 1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b15_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
