cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boxqp.py tests/test_gpu_round2.py -q -x -k "cfg5 or 124 or edge or inside" 2>&1 | tail -2
timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu --batch 524288 > gpurun_out/b43_cfg5.json 2> gpurun_out/b43_cfg5.err
python -c "
import json
d=json.loads(open('gpurun_out/b43_cfg5.json').read().strip().splitlines()[-1]); print('cfg5 2^19', round(d['ms_per_step'],1), d['value'], (d.get('solved_only') or {}).get('value'))
"
