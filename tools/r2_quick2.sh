#!/bin/bash
# quick GPU check of the fp32-grade (tc32) chain: parity tests that touch it + bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc32 or fp32 or vision or f16 or full_size or net" > gpurun_out/quick2_tests.log 2>&1
tail -4 gpurun_out/quick2_tests.log
for net in tc32 bf16; do
  timeout 300 python bench.py --net $net --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/quick2_$net.json 2> gpurun_out/quick2_$net.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/quick2_$net.json").read().strip().splitlines()[-1])
    print("$net", "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "net us %.2f" % d["roofline"]["avg_launch_us"], "frac %.4f" % d["roofline"]["frac"])
except Exception as e:
    print("$net failed", e)
PY
done
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/quick2_cfg4.json 2>/dev/null
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/quick2_cfg5.json 2>/dev/null
python - <<PY
import json
for w in ("cfg4", "cfg5"):
    try:
        d = json.loads(open("gpurun_out/quick2_%s.json" % w).read().strip().splitlines()[-1])
        print(w, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"])
    except Exception as e:
        print(w, "failed", e)
PY
