python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large_pippenger" 2>&1 | tail -2
python tools/msm_latency.py 16 22 2>&1 | tail -7
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ba_launches.csv python tools/msm_latency.py 22 22 > gpurun_out/ba_lat.log 2>&1
