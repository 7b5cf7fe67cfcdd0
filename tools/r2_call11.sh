#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c11_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/c11_parity.json 2>/dev/null
timeout 200 python tools/stress.py 60 > gpurun_out/c11_stress.log 2>&1; echo "rc=$?" >> gpurun_out/c11_stress.log
timeout 900 python bench.py --no-competitors > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
for f in gpurun_out/c11_tests.log gpurun_out/c11_stress.log; do echo "## $f: $(tail -3 $f | tr '\n' ' ' | cut -c1-300)"; done
grep -n "FAILED\|Error\|error" gpurun_out/c11_tests.log | head -20
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c11_bench.json").read().strip().splitlines()[-1])
print("decode", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "prefill", round(d["prefill"]["tflops"],1))
print("act-order", json.dumps(d["extra"]["act_order_8b"])[:400])
PY
