python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large_pippenger" 2>&1 | tail -2
python tools/msm_latency.py 20 22 2>&1 | tail -3
CDP_BA_FINISH_ROUNDS=0 python tools/msm_latency.py 20 22 2>&1 | tail -3
CDP_BA_FINISH_ROUNDS=5 python tools/msm_latency.py 20 22 2>&1 | tail -3
CDP_BIG_BA_MIN_LOG2=16 python tools/msm_latency.py 16 19 2>&1 | tail -4
