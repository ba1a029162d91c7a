python -m pytest tests/test_gpu_parity.py tests/test_gpu_fixed.py -x -q -m gpu -k "large_pippenger or tree" 2>&1 | tail -2
python tools/msm_latency.py 20 22 2>&1 | tail -3
python tools/fixed_bench.py 2>&1 | tail -10
python tools/prover_timing.py 252 4096 8 2>&1 | tail -2 | cut -c1-130
python tools/prover_timing.py 252 4096 8 2>&1 | tail -2 | cut -c1-130
