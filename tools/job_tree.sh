python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large_pippenger" 2>&1 | tail -2
python tools/msm_latency.py 20 22 2>&1 | tail -3
echo "c=17 at 2^20"; CDP_BIG_C=17 python tools/msm_latency.py 20 20 2>&1 | tail -1
echo "c=18 at 2^21"; CDP_BIG_C=18 python tools/msm_latency.py 21 21 2>&1 | tail -1
echo "c=16 at 2^19 BA"; CDP_BIG_BA_MIN_LOG2=16 CDP_BIG_C=16 python tools/msm_latency.py 19 19 2>&1 | tail -1
echo "c=15 at 2^19 BA"; CDP_BIG_BA_MIN_LOG2=16 python tools/msm_latency.py 17 19 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ba_launches.csv python tools/msm_latency.py 22 22 > gpurun_out/ba_lat.log 2>&1
