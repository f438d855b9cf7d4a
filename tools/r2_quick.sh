#!/bin/bash
# quick GPU check of a kernel change: variant parity tests + bench of cfg2 with and without the change
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or full_search or fidelity or philox" > gpurun_out/quick_tests.log 2>&1
tail -5 gpurun_out/quick_tests.log
for v in 1 0; do
  SMZ_M32=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/quick_m32_$v.json 2> gpurun_out/quick_m32_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/quick_m32_$v.json").read().strip().splitlines()[-1])
    print("SMZ_M32=$v", "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "net us %.2f" % d["roofline"]["avg_launch_us"], "frac %.4f" % d["roofline"]["frac"])
except Exception as e:
    print("SMZ_M32=$v failed", e)
PY
done
SMZ_BF16_TIMELINE=1 timeout 120 python tools/diag_timeline.py 2>&1 | tail -20
