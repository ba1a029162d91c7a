# Round-2 final captures (run under gpurun on one B200): the dominant kernels with the final code (XYZZ accumulators, load-ordered slots).
set -x
cd $GRAFT_REPO_ROOT
PIPES=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__cycles_active.avg,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fixed_msm --launch-skip 44 --launch-count 1 -f -o gpurun_out/r02b_fixed python tools/prover_profile.py 252 1024 1 > gpurun_out/r02b_ncu_f.log 2>&1
timeout 400 ncu --clock-control none -k regex:k_fixed_msm --launch-skip 44 --launch-count 1 --metrics $PIPES --csv --log-file gpurun_out/r02b_fixed_pipes.csv python tools/prover_profile.py 252 1024 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_big_accumulate --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02b_bigacc python tools/msm_latency.py 22 22 > gpurun_out/r02b_ncu_a.log 2>&1
timeout 400 ncu --clock-control none -k regex:k_big_accumulate --launch-skip 2 --launch-count 1 --metrics $PIPES --csv --log-file gpurun_out/r02b_bigacc_pipes.csv python tools/msm_latency.py 22 22 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_msm_buckets --launch-skip 8 --launch-count 1 -f -o gpurun_out/r02b_buckets python tools/prover_profile.py 252 1024 1 > gpurun_out/r02b_ncu_b.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_prover_B512_lane1.csv python tools/prover_profile.py 252 512 1 > /dev/null 2>&1
ls -la gpurun_out/r02b_*
