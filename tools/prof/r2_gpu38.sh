cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -c 1 -s 1 -f"
PYTHONPATH=tests timeout 300 python tools/prof/prof_coop.py && PYTHONPATH=tests timeout 900 $N -k regex:boxqp_ipm_coop -o gpurun_out/r02_coop_stage_major python tools/prof/prof_coop.py > gpurun_out/ncu_d.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_coop_stage_major.ncu-rep > gpurun_out/r02_coop_stage_major.ncu.txt 2>&1
cat gpurun_out/r02_coop_stage_major.ncu.txt
