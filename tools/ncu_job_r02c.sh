# Round-2 captures of the batched-affine kernels (run under gpurun on one B200).
set -x
cd $GRAFT_REPO_ROOT
PIPES=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,sm__cycles_active.avg,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio
# round 1 (gather) and round 2 (array) of the third 2^22 MSM msm_latency.py runs: 11 rounds per MSM
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ba_round --launch-skip 22 --launch-count 2 -f -o gpurun_out/r02c_ba_round python tools/msm_latency.py 22 22 > gpurun_out/r02c_ncu_a.log 2>&1
timeout 500 ncu --clock-control none -k regex:k_ba_round --launch-skip 22 --launch-count 2 --metrics $PIPES --csv --log-file gpurun_out/r02c_ba_round_pipes.csv python tools/msm_latency.py 22 22 > /dev/null 2>&1
# the tree's first round and its finishing lane kernel in the isolated fixed-base bench (B=1024, K=4, n=128)
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fixed_ba_first --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02c_fixed_first python tools/fixed_bench.py > gpurun_out/r02c_ncu_f.log 2>&1
timeout 500 ncu --clock-control none -k regex:k_fixed_ba --launch-skip 15 --launch-count 5 --metrics $PIPES --csv --log-file gpurun_out/r02c_fixed_tree_pipes.csv python tools/fixed_bench.py > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches_prover_B512_lane1.csv python tools/prover_profile.py 252 512 1 > /dev/null 2>&1
ls -la gpurun_out/r02c_*
