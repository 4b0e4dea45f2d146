set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -c 1 -s 1 -f"
timeout 300 python tools/prof/prof_boxqp_cfg3.py 262144 && timeout 900 $N -k regex:boxqp_ipm -o gpurun_out/r02_boxqp_cfg3_staged3 python tools/prof/prof_boxqp_cfg3.py 262144 > gpurun_out/ncu_a.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_boxqp_cfg3_staged3.ncu-rep > gpurun_out/r02_boxqp_cfg3_staged3.ncu.txt 2>&1
cat gpurun_out/r02_boxqp_cfg3_staged3.ncu.txt
