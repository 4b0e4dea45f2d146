set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; wl=$2; shift; shift; e=$1; shift; env $e timeout 600 python bench.py "$@" --workload $wl --steps 2 --warmup 1 --no-cpu > gpurun_out/b14_$name.json 2> gpurun_out/b14_$name.err; }
run cfg4_32768 cfg4 X=1 --batch 32768
run cfg4_37888 cfg4 X=1 --batch 37888
run cfg4_75776 cfg4 X=1 --batch 75776
run cfg4_65536 cfg4 X=1
run obstacle obstacle X=1 --steps 2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_nocpu.json 2> gpurun_out/r02_bench_nocpu.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/b14_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], round(d['ms_per_step'],2), d['clocks'].get('power_w'))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
P
