# Round-2 ncu evidence: launch list of the default bench command + one --set full capture per dominant kernel.
# Run under gpurun AFTER the same commands have exited 0 without ncu; copy the digests into profiles/.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -c 1 -s 1 -f"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_nocpu.json 2> gpurun_out/r02_bench_nocpu.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 300 python tools/prof/prof_boxqp_cfg3.py && timeout 900 $N -k regex:boxqp_ipm -o gpurun_out/r02_boxqp_cfg3_final python tools/prof/prof_boxqp_cfg3.py > gpurun_out/ncu_a.log 2>&1
timeout 300 python tools/prof/prof_boxqp_cfg3.py 151552 f32 && timeout 900 $N -k regex:boxqp_ipm -o gpurun_out/r02_boxqp_cfg3_f32 python tools/prof/prof_boxqp_cfg3.py 151552 f32 > gpurun_out/ncu_b.log 2>&1
timeout 300 python tools/prof/prof_rti.py && timeout 1500 $N -k regex:rti_closed_loop -o gpurun_out/r02_rti_final python tools/prof/prof_rti.py > gpurun_out/ncu_c.log 2>&1
PYTHONPATH=tests timeout 300 python tools/prof/prof_coop.py && PYTHONPATH=tests timeout 900 $N -k regex:boxqp_ipm_coop -o gpurun_out/r02_coop_dmma python tools/prof/prof_coop.py > gpurun_out/ncu_d.log 2>&1
timeout 300 python tools/prof/prof_k1.py && timeout 900 $N -k regex:riccati_reg -o gpurun_out/r02_k1 python tools/prof/prof_k1.py > gpurun_out/ncu_e.log 2>&1
timeout 300 python tools/prof/prof_lq_solve.py && timeout 900 $N -k regex:lq_solve_krylov -o gpurun_out/r02_lq_krylov python tools/prof/prof_lq_solve.py > gpurun_out/ncu_f.log 2>&1
for r in r02_boxqp_cfg3_final r02_boxqp_cfg3_f32 r02_rti_final r02_coop_dmma r02_k1 r02_lq_krylov; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/$r.ncu.txt 2>&1
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    for k in ('sm__inst_executed_pipe_tensor_op_dmma.sum','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.sum','smsp__inst_executed.sum'):
        if k in h: print(' ',k, r[h.index(k)])
" >> gpurun_out/$r.ncu.txt
  tail -40 gpurun_out/$r.ncu.txt | head -30
done
rm -f gpurun_out/*.ncu-rep.tmp
ls -la gpurun_out/*.ncu-rep
