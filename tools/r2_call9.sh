#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c9_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/c9_parity.json 2>/dev/null
timeout 200 python tools/stress.py 100 > gpurun_out/c9_stress.log 2>&1; echo "rc=$?" >> gpurun_out/c9_stress.log
timeout 120 python tools/microbench.py gemm 2048 > gpurun_out/c9_gemm_bench.log 2>&1
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/c9_bench_ref.json 2> gpurun_out/c9_bench_ref.err
tail -3 gpurun_out/c9_tests.log; tail -2 gpurun_out/c9_stress.log
grep -E "GEMM M|cuBLAS" gpurun_out/c9_gemm_bench.log
grep -E "^\[bench|Elapsed" gpurun_out/c9_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c9_bench.json").read().strip().splitlines()[-1])
print("decode", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "prefill", round(d["prefill"]["tflops"],1), round(d["prefill"]["roofline"]["frac"],3))
print("cpu_baseline", d.get("cpu_baseline"))
print(json.dumps(d.get("competitors"))[:2500])
r = json.loads(open("gpurun_out/c9_bench_ref.json").read().strip().splitlines()[-1])
print("ref arm", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample"][:300])
PY
