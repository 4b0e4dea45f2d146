cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -c 1 -s 1 -f"
for st in f64 mix; do
MPC_QP_STORE=$st timeout 300 python tools/prof/prof_rti_small.py && MPC_QP_STORE=$st timeout 900 $N -k regex:rti_closed_loop -o gpurun_out/r02_rti_small_$st python tools/prof/prof_rti_small.py > gpurun_out/ncu_c.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_rti_small_$st.ncu-rep > gpurun_out/r02_rti_small_$st.ncu.txt 2>&1
cat gpurun_out/r02_rti_small_$st.ncu.txt
done
