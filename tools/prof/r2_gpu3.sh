set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest3.log
tail -15 gpurun_out/pytest3.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --workload ${WL:-cfg3} --steps 3 --warmup 3 --no-cpu > gpurun_out/b3_$name.json 2> gpurun_out/b3_$name.err; }
run cfg3_default X=1
run cfg3_refill0 MPC_QP_REFILL=0
run cfg3_refill4 MPC_QP_REFILL=4
run cfg3_refill16 MPC_QP_REFILL=16
run cfg3_refill8_minb3 MPC_QP_MINB=3
run cfg3_refill8_pf0 MPC_QP_PREFETCH=0
run cfg3_refill8_pf2 MPC_QP_PREFETCH=2
run cfg3_refill0_minb3 MPC_QP_REFILL=0 MPC_QP_MINB=3
WL=cfg4 run cfg4 X=1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b3_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['fp_pipe']['mean_iters_per_solve'], d['summary'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
P
