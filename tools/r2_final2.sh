#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/fin2_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/fin2_parity.json 2>/dev/null
timeout 120 python tools/san_midm_graph.py 4096 4096 128 24 > gpurun_out/fin2_chain.log 2>&1; echo "rc=$?" >> gpurun_out/fin2_chain.log
timeout 250 python tools/microbench.py midm 16 64 128 > gpurun_out/fin2_midm_bench.log 2>&1
timeout 300 python tools/stress.py 150 > gpurun_out/fin2_stress.log 2>&1; echo "rc=$?" >> gpurun_out/fin2_stress.log
timeout 900 python bench.py > gpurun_out/fin2_bench.json 2> gpurun_out/fin2_bench.err
timeout 300 python bench.py --impl reference > gpurun_out/fin2_bench_ref.json 2> gpurun_out/fin2_bench_ref.err
for f in gpurun_out/fin2_tests.log gpurun_out/fin2_chain.log gpurun_out/fin2_stress.log; do echo "## $f: $(tail -2 $f | tr '\n' ' ' | cut -c1-200)"; done
grep "MIDM bits=4 g=128\|MIDM bits=8" gpurun_out/fin2_midm_bench.log | cut -c1-210
python - <<'PY'
import json
d = json.loads(open("gpurun_out/fin2_bench.json").read().strip().splitlines()[-1])
print("decode", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "prefill", round(d["prefill"]["tflops"],1), round(d["prefill"]["roofline"]["frac"],3), d["clocks"])
print(json.dumps(d["competitors"]["per_shape_us"])[:1600])
r = json.loads(open("gpurun_out/fin2_bench_ref.json").read().strip().splitlines()[-1])
print("ref arm", r["value"], r["cpu_baseline"]["cores"], "in-arm cpu", d["cpu_baseline"]["value"])
PY
