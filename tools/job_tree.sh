for st in 0 1000 3000 8000; do CDP_LANE_STAGGER_US=$st python bench.py --no-cpu-baseline --msm-sizes '' --no-extras --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stagger=$st proofs/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],1))"; done
