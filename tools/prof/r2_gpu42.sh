cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_boxqp.py tests/test_gpu_rti.py -q -x -k "obstacle or rows or fused" 2>&1 | tail -2
timeout 300 python bench.py --workload obstacle --steps 1 --warmup 1 --no-cpu > gpurun_out/b42_obst.json 2>gpurun_out/b42_obst.err
python -c "
import json
d=json.loads(open('gpurun_out/b42_obst.json').read().strip().splitlines()[-1]); print('obstacle', round(d['ms_per_step'],1), d['value'])
"
