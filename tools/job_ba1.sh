python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large_pippenger or linearity" 2>&1 | tail -2
python tools/msm_latency.py 13 22 2>&1 | tail -10
CDP_BIG_MIN_LOG2=13 python tools/msm_latency.py 13 16 2>&1 | tail -4
for t in 56832 227328; do echo "THREADS=$t"; CDP_BA_THREADS=$t python tools/msm_latency.py 18 22 2>&1 | tail -5; done
for k in 32 128; do echo "KMAX=$k"; CDP_BA_KMAX=$k python tools/msm_latency.py 22 22 2>&1 | tail -1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ba_launches.csv python tools/msm_latency.py 22 22 > gpurun_out/ba_lat.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ba_launches18.csv python tools/msm_latency.py 18 18 > gpurun_out/ba_lat18.log 2>&1
