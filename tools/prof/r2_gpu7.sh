set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/prof/prof_rti.py && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rti_closed_loop -c 1 -s 1 -o gpurun_out/r02_rti -f python tools/prof/prof_rti.py > gpurun_out/ncu_rti.log 2>&1
tail -3 gpurun_out/ncu_rti.log
python tools/ncu_summary.py gpurun_out/r02_rti.ncu-rep > gpurun_out/r02_rti.ncu.txt 2>&1; cat gpurun_out/r02_rti.ncu.txt
ncu -i gpurun_out/r02_rti.ncu-rep --page source --csv > gpurun_out/r02_rti.source.csv 2>/dev/null; wc -l gpurun_out/r02_rti.source.csv
timeout 300 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu > gpurun_out/b7_cfg5_dmma.json 2> gpurun_out/b7_cfg5_dmma.err
MPC_COOP_DMMA=0 timeout 300 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu > gpurun_out/b7_cfg5_scalar.json 2> gpurun_out/b7_cfg5_scalar.err
MPC_COOP_MINB=2 timeout 300 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu > gpurun_out/b7_cfg5_dmma_m2.json 2> gpurun_out/b7_cfg5_dmma_m2.err
MPC_COOP_MINB=3 timeout 300 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu > gpurun_out/b7_cfg5_dmma_m3.json 2> gpurun_out/b7_cfg5_dmma_m3.err
timeout 600 python -m pytest tests/test_gpu_boxqp.py -q -x -k "cfg5 or 124 or edge" > gpurun_out/pytest7.log 2>&1; tail -3 gpurun_out/pytest7.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b7_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d.get('solved_only'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
