set -x
cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_big_accumulate|k_decompress" --launch-skip 4 --launch-count 2 -f -o gpurun_out/r01_verifier_kernels python tools/verifier_profile.py 252 4096 1 > gpurun_out/s41_ncu_v.log 2>&1
tail -3 gpurun_out/s41_ncu_v.log
ls -la gpurun_out/r01_verifier_kernels.ncu-rep
