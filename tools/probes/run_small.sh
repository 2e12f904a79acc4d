python - <<'PY' > gpurun_out/small_case.log 2>&1
import json, sys
sys.path.insert(0, '.')
from tests.gpu_cases import CASES
print(json.dumps(CASES["small_batch_step"](), default=str))
PY
tail -c 1500 gpurun_out/small_case.log
python tools/profile_small.py large-v3 1 40 > gpurun_out/ps_small_b1_v3.json 2> gpurun_out/ps_small_b1_v3.err; tail -3 gpurun_out/ps_small_b1_v3.err
python tools/time_exact.py > gpurun_out/time_exact_v3.log 2>&1; tail -8 gpurun_out/time_exact_v3.log
