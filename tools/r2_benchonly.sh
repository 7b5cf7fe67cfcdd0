#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
timeout 300 python bench.py --impl reference > gpurun_out/c16_bench_ref.json 2> gpurun_out/c16_bench_ref.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c16_bench.json").read().strip().splitlines()[-1])
print("decode", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "prefill", round(d["prefill"]["tflops"],1), d["clocks"])
print("mixtral", d["extra"]["mixtral_8x7b_tp4"]["decode_tok_s"], "act", d["extra"]["act_order_8b"]["decode_tok_s"])
r = json.loads(open("gpurun_out/c16_bench_ref.json").read().strip().splitlines()[-1])
print("ref arm", r["value"], r["cpu_baseline"]["cores"])
PY
