set -x
cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_msm_buckets<5" --launch-skip 2 --launch-count 1 -f -o gpurun_out/r01_buckets_v5 python tools/prover_profile.py 252 1024 1 > gpurun_out/s24_ncu_b.log 2>&1
python bench.py > gpurun_out/s24_bench.json 2> gpurun_out/s24_bench.err
python bench.py --impl reference > gpurun_out/s24_ref.json 2> gpurun_out/s24_ref.err
cut -c1-330 gpurun_out/s24_bench.json; cut -c1-200 gpurun_out/s24_ref.json
