# cfg4 (fused RTI loop, (4,2) stages) against the L2 prefetch distance of the workspace rows
for pf in 0 1 2 3 4 6; do echo -n "cfg4 pf=$pf: "; MPC_QP_PREFETCH=$pf python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['summary']['sum_iters'])"; done
