for v in 1 0; do B200W_CROSS_STREAM=$v python tools/time_cross.py; done
python - <<'PY'
import json, sys
sys.path.insert(0, '.')
from tests.gpu_cases import CASES
for n in ("decoder_attention", "splitk_decode_ops", "config3_small_batch64"):
    print(n, json.dumps(CASES[n](), default=str)[:600])
PY
for v in 1 0; do B200W_CROSS_STREAM=$v python tools/profile_step.py --skip-encoder; done
