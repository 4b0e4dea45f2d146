cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all_b30.log 2>&1; tail -5 gpurun_out/pytest_all_b30.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
