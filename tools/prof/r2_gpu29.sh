set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -c 1 -s 1 -f"
timeout 300 python tools/prof/prof_rti_steady.py && timeout 1500 $N -k regex:rti_closed_loop -o gpurun_out/r02_rti_steady python tools/prof/prof_rti_steady.py > gpurun_out/ncu_c.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_rti_steady.ncu-rep > gpurun_out/r02_rti_steady.ncu.txt 2>&1
cat gpurun_out/r02_rti_steady.ncu.txt
