set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --steps 2 --warmup 2 --no-cpu > gpurun_out/b10_$name.json 2> gpurun_out/b10_$name.err; }
run cfg3_m4 cfg3 X=1
run cfg3_m3 cfg3 MPC_QP_MINB=3
run cfg3_m4_pf2 cfg3 MPC_QP_PREFETCH=2
run cfg3_f32 cfg3 X=1 --dtype f32
run cfg3_N20 cfg3 X=1 --horizon 20
run cfg4 cfg4 X=1
run cfg4_f64 cfg4 MPC_QP_STORE=f64
run cfg4_f32 cfg4 X=1 --dtype f32
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b10_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'), d['summary'].get('n_infeasible'), d['summary'].get('sum_iters'), d['summary'].get('sum_saturated'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest10.log
tail -25 gpurun_out/pytest10.log
