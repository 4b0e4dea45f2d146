cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 1 11 12 13 14; do
echo "== variant $v"
MPC_QP_PREFETCH=$v timeout 300 python -m pytest tests/test_gpu_boxqp.py -q -x -k "full_size_properties_cfg3" 2>&1 | tail -2
done
MPC_QP_STAGED=0 timeout 300 python -m pytest tests/test_gpu_boxqp.py -q -x -k "full_size_properties_cfg3" 2>&1 | tail -2
