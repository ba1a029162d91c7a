for w in 1 0 1 0; do CDP_TRANSCRIPT_WARP=$w python bench.py --no-cpu-baseline --msm-sizes '' --no-extras --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warp=$w proofs/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'verifies/s', round(d['verify']['value']), 'ms', d['verify']['ms_per_step'])"; done
