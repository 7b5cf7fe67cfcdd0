#!/bin/bash
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; ( "$@" ) > gpurun_out/c8_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/c8_$tag.log; echo "## $tag: $(grep -E '^ok|ok$|rc=' gpurun_out/c8_$tag.log | tail -4 | tr '\n' ' ')"; }
run full_c36 timeout 120 python tools/san_midm_graph.py 4096 4096 16 36
run m128_k14336 timeout 120 python tools/san_midm_graph.py 14336 4096 128 12
timeout 300 python tools/san_midm.py > gpurun_out/c8_san_midm.log 2>&1
timeout 300 python tools/san_moe.py > gpurun_out/c8_san_moe.log 2>&1
run stress timeout 400 python tools/stress.py 300
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c8_default_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/c8_parity_default.json 2>/dev/null
timeout 200 python tools/microbench.py midm 16 128 > gpurun_out/c8_midm_bench.log 2>&1
timeout 120 python tools/microbench.py gemm 2048 > gpurun_out/c8_gemm_bench.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-competitors --no-cpu-baseline > gpurun_out/c8_bench_extra.json 2> gpurun_out/c8_bench_extra.err
for f in gpurun_out/c8_san*.log gpurun_out/c8_*tests.log; do echo "## $f: $(tail -1 $f | cut -c1-220)"; done
grep MIDM gpurun_out/c8_midm_bench.log | head -30
grep -E "GEMM M|cuBLAS" gpurun_out/c8_gemm_bench.log
grep "^\[bench" gpurun_out/c8_bench_extra.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c8_bench_extra.json").read().strip().splitlines()[-1])
    print("decode", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "prefill", round(d["prefill"]["tflops"],1))
    print(json.dumps(d.get("extra"), indent=None)[:4000])
except Exception as e:
    print("unreadable:", e)
PY
