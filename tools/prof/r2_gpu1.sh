set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest1.log
tail -5 gpurun_out/pytest1.log
for wl in cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/b1_$wl.json 2> gpurun_out/b1_$wl.err; echo "$wl rc=$?"
done
MPC_QP_STORE=f64 timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/b1_cfg3_f64store.json 2>&1
for mb in 3 6; do MPC_QP_MINB=$mb timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/b1_cfg3_minb$mb.json 2>&1; done
for pf in 0 1 3; do MPC_QP_PREFETCH=$pf timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/b1_cfg3_pf$pf.json 2>&1; done
MPC_QP_PREFETCH=2 timeout 600 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu > gpurun_out/b1_cfg4_pf2.json 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b1_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['fp_pipe']['mean_iters_per_solve'], d['summary'])
    except Exception as e: print(f, 'ERR', e)
P
