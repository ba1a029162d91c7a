python -m pytest tests/test_gpu_prover.py tests/test_gpu_verifier.py tests/test_gpu_transcript.py -x -q -m gpu 2>&1 | tail -2
python tools/prover_profile.py 252 512 1 2>&1 | grep "prove_stage\|sum\|ell=" | cut -c1-120
python tools/prover_timing.py 252 128 4 2>&1 | grep "B=128" | cut -c1-150
python tools/verifier_profile.py 252 512 1 2>&1 | grep "transcript\|sum\|ell=" | cut -c1-150
