set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest2.log
tail -15 gpurun_out/pytest2.log
timeout 600 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu > gpurun_out/b2_cfg4.json 2> gpurun_out/b2_cfg4.err; echo "cfg4 rc=$?"; tail -3 gpurun_out/b2_cfg4.err
timeout 300 python tools/prof/prof_boxqp_cfg3.py && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:boxqp_ipm -c 1 -s 1 -o gpurun_out/r02_boxqp_cfg3 -f python tools/prof/prof_boxqp_cfg3.py > gpurun_out/ncu_cfg3.log 2>&1
tail -3 gpurun_out/ncu_cfg3.log
python tools/ncu_summary.py gpurun_out/r02_boxqp_cfg3.ncu-rep > gpurun_out/r02_boxqp_cfg3.ncu.txt 2>&1; cat gpurun_out/r02_boxqp_cfg3.ncu.txt
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], d['ms_per_step'], d['roofline']['fp_pipe']['mean_iters_per_solve'], d['summary'])
    except Exception as e: print(f, 'ERR', e)
P
