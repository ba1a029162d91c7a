cd $GRAFT_REPO_ROOT
echo "== msm latency default"; python tools/msm_latency.py 8 18 2>&1 | tail -13
echo "== chunk 32"; CDP_MSM_SINGLE_CHUNK=32 python tools/msm_latency.py 10 14 2>&1 | tail -6
echo "== chunk 128"; CDP_MSM_SINGLE_CHUNK=128 python tools/msm_latency.py 10 16 2>&1 | tail -8
echo "== big from 2^14"; CDP_BIG_MIN_LOG2=14 python tools/msm_latency.py 14 18 2>&1 | tail -6
echo "== fixed default"; python tools/fixed_bench.py 2>&1 | tail -6
echo "== fixed bulk"; CDP_FIXED_BULK=1 python tools/fixed_bench.py 2>&1 | tail -6
echo "== fixed bulk occ4"; CDP_FIXED_BULK=1 CDP_OCC_FIXED=4 python tools/fixed_bench.py 2>&1 | tail -6
echo "== parity with bulk"; CDP_FIXED_BULK=1 python -m pytest tests/test_gpu_fixed.py tests/test_gpu_prover.py -x -q -m gpu -k "not 1020" 2>&1 | tail -3
echo "== prover lanes at B=128"; for l in 2 4 8; do python tools/prover_timing.py 252 128 $l 2>&1 | grep "B=128"; done
