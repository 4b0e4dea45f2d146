set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=model_predictive_control_b200/lib/variants
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/b5_$name.json 2> gpurun_out/b5_$name.err; }
for v in v0_gpu1 v2_stage v3_ptrs v4_stage_ptrs; do
  run ${v}_cfg3_m4 cfg3 MPC_B200_LIB=$V/$v.so MPC_QP_REFILL=0 MPC_QP_PREFETCH=1
  run ${v}_cfg3_m3 cfg3 MPC_B200_LIB=$V/$v.so MPC_QP_REFILL=0 MPC_QP_PREFETCH=1 MPC_QP_MINB=3
  run ${v}_cfg4 cfg4 MPC_B200_LIB=$V/$v.so
done
run main_cfg3_m4 cfg3 MPC_QP_REFILL=0 MPC_QP_PREFETCH=1
run main_cfg3_m3 cfg3 MPC_QP_REFILL=0 MPC_QP_PREFETCH=1 MPC_QP_MINB=3
run v4_cfg3_refill8_m3 cfg3 MPC_B200_LIB=$V/v4_stage_ptrs.so MPC_QP_REFILL=8 MPC_QP_MINB=3
run v4_cfg3_refill8_m4 cfg3 MPC_B200_LIB=$V/v4_stage_ptrs.so MPC_QP_REFILL=8
run v2_cfg3_refill8_m3 cfg3 MPC_B200_LIB=$V/v2_stage.so MPC_QP_REFILL=8 MPC_QP_MINB=3
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b5_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
P
