# quick cfg3 / cfg4 timing (no CPU leg): value, ms per step, total iterations, summed cost
for w in cfg3 cfg4; do st=5; [ $w = cfg4 ] && st=2; python bench.py --workload $w --steps $st --warmup 3 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['ms_per_step'], d['summary']['sum_iters'], d['summary']['sum_cost'])"; done
