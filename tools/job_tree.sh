python -m pytest tests/test_gpu_transcript.py tests/test_gpu_prover.py tests/test_gpu_verifier.py -x -q -m gpu 2>&1 | tail -3
python tools/prover_profile.py 252 512 1 2>&1 | grep "prove_stage\|transcript\|sum\|ell=" | cut -c1-120
python tools/verifier_profile.py 252 512 1 2>&1 | grep "transcript\|sum\|ell=" | cut -c1-150
python tools/verifier_profile.py 252 4096 8 2>&1 | grep "transcript\|sum\|ell=" | cut -c1-150
python tools/prover_timing.py 252 128 4 2>&1 | grep "B=128" | cut -c1-150
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vlaunches.csv python tools/verifier_profile.py 252 512 1 > /dev/null 2>&1
