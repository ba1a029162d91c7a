set -x
cd $GRAFT_REPO_ROOT
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_launches_v3.csv python bench.py --steps 1 --warmup 1 --batch 1024 --lanes 1 --no-cpu-baseline > gpurun_out/s22_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msm_buckets --launch-skip 20 --launch-count 1 -f -o gpurun_out/r01_buckets_v5 python tools/prover_profile.py 252 1024 1 > gpurun_out/s22_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_smul_jobs --launch-skip 14 --launch-count 1 -f -o gpurun_out/r01_smul_v1 python tools/prover_profile.py 252 1024 1 > gpurun_out/s22_ncu_s.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fixed_msm --launch-skip 44 --launch-count 1 -f -o gpurun_out/r01_fixed_v3 python tools/prover_profile.py 252 1024 1 > gpurun_out/s22_ncu_f.log 2>&1
CDP_VERIFY_TRACE=1 python tools/verifier_profile.py 252 1024 1 > gpurun_out/s22_vtrace.txt 2>&1
tail -25 gpurun_out/s22_vtrace.txt
ls -la gpurun_out/*.ncu-rep
