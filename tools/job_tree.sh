for v in 20 19 18 20 19 18; do CDP_BIG_BA_MIN_LOG2=$v python bench.py --no-cpu-baseline --msm-sizes '' --no-extras --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ba_min=2^$v proofs/s', round(d['value']), 'verifies', round(d['verify']['value']), 'ms', round(d['verify']['ms_per_step'],1), 'worst', round(d['verify']['worst_case']['value']) if 'worst_case' in d['verify'] else '-')"; done
