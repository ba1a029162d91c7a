# Round-2 profiling job (run under gpurun on one B200): ncu captures of the dominant kernels with the final code, plus the metric names
# this ncu knows for the FMA sub-pipes.  Summaries go to profiles/ (tools/ncu_summary.py).
set -x
cd $GRAFT_REPO_ROOT
ncu --query-metrics 2>/dev/null | grep -iE "fma|imad|pipe_alu|inst_executed_op" > gpurun_out/r02_ncu_metric_names.txt
wc -l gpurun_out/r02_ncu_metric_names.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fixed_msm --launch-skip 44 --launch-count 1 -f -o gpurun_out/r02_fixed python tools/prover_profile.py 252 1024 1 > gpurun_out/r02_ncu_f.log 2>&1
timeout 400 ncu --clock-control none -k regex:k_fixed_msm --launch-skip 44 --launch-count 1 --metrics sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__cycles_active.avg,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02_fixed_pipes.csv python tools/prover_profile.py 252 1024 1 > gpurun_out/r02_ncu_f2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_big_accumulate --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_bigacc python tools/msm_latency.py 20 20 > gpurun_out/r02_ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_prove_stage --launch-skip 54 --launch-count 1 -f -o gpurun_out/r02_stage python tools/prover_profile.py 252 1024 1 > gpurun_out/r02_ncu_s.log 2>&1
ls -la gpurun_out/*.ncu-rep
