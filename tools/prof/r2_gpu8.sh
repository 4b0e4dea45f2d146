set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=model_predictive_control_b200/lib/variants
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 2 --warmup 2 --no-cpu > gpurun_out/b8_$name.json 2> gpurun_out/b8_$name.err; }
run main_cfg3_m4 cfg3 X=1
run main_cfg3_m3 cfg3 MPC_QP_MINB=3
run v4_cfg3_m4 cfg3 MPC_B200_LIB=$V/v4_stage_ptrs.so
run v4_cfg3_m3 cfg3 MPC_B200_LIB=$V/v4_stage_ptrs.so MPC_QP_MINB=3
run v3_cfg3_m4 cfg3 MPC_B200_LIB=$V/v3_ptrs.so
run v2_cfg3_m4 cfg3 MPC_B200_LIB=$V/v2_stage.so
run main_cfg4 cfg4 X=1
run main_cfg4_pf1 cfg4 MPC_QP_PREFETCH=1
run main_cfg4_pf2 cfg4 MPC_QP_PREFETCH=2
run main_cfg4_f64 cfg4 MPC_QP_STORE=f64
run main_cfg4_f64_pf1 cfg4 MPC_QP_STORE=f64 MPC_QP_PREFETCH=1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
