set -x
cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_msm_buckets --launch-skip 20 --launch-count 10 -f -o gpurun_out/r01_buckets_v5 python tools/prover_profile.py 252 1024 1 > gpurun_out/s25_ncu_b.log 2>&1
python tools/msm_sweep.py 22 > gpurun_out/s25_sweep.txt 2>&1
tail -4 gpurun_out/s25_sweep.txt
