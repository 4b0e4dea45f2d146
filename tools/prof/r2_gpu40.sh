cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in a74df4c 35bf2bb bcb674f HEAD; do
  if [ $c = HEAD ]; then L=""; else L="MPC_B200_LIB=$PWD/model_predictive_control_b200/lib/variants/bisect_$c.so"; fi
  env $L timeout 300 python bench.py --workload obstacle --steps 1 --warmup 1 --no-cpu > gpurun_out/b40_obst_$c.json 2> gpurun_out/b40_obst_$c.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/b40_obst_$c.json').read().strip().splitlines()[-1]); print('$c', round(d['ms_per_step'],1))
except Exception as e: print('$c','ERR', open('gpurun_out/b40_obst_$c.err').read()[-400:])
P
done
