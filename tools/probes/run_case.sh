python - <<'PY'
import json, sys
sys.path.insert(0, '.')
from tests.gpu_cases import CASES
for n in sys.argv[1:] or ("ring_attention_edges", "small_batch_step"):
    print(n, json.dumps(CASES[n](), default=str)[:900])
PY
python tools/probes/run_small_ab.py 2>&1 | tail -1
