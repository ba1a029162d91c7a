python -m pytest tests/test_gpu_prover.py -x -q -m gpu 2>&1 | tail -2
python tools/prover_profile.py 252 512 1 2>&1 | grep "msm_fixed\|sum\|ell=" | cut -c1-130
for i in 1 2; do python bench.py --no-cpu-baseline --msm-sizes '' --no-extras --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('proofs/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'])"; done
