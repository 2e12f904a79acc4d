python - <<'PY'
import json, sys
sys.path.insert(0, '.')
from tests.gpu_cases import CASES
for n in ("filter_argmax", "sampling_distribution", "decode_tiny"):
    print(n, json.dumps(CASES[n](), default=str)[:400])
PY
python tools/profile_step.py --skip-encoder --eager 2 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print({k: round(v['avg_us'],1) for k,v in d['eager_kernels_ms_per_step'].items()})"
python tools/profile_step.py --skip-encoder 2>&1 | tail -1
python tools/probes/run_small_ab.py 2>&1 | tail -1
