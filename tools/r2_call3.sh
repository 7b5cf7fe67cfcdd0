#!/bin/bash
# round-2 third GPU call: k-block-parallel dequant groups in the small-batch tier, grouped MoE kernels (first contact),
# hybrid decode dispatch + unrolled activation staging
set -u
mkdir -p gpurun_out
timeout 300 python tools/san_midm.py > gpurun_out/c3_san_midm.log 2>&1
timeout 300 python tools/san_moe.py > gpurun_out/c3_san_moe.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python tools/san_moe.py > gpurun_out/c3_san_moe_memcheck.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c3_default_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/c3_parity_default.json 2>/dev/null
timeout 300 python tools/microbench.py midm 16 64 128 > gpurun_out/c3_midm_bench.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c3_bench_hybrid.json 2> gpurun_out/c3_bench_hybrid.err
B2Q_DECODE_V2=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c3_bench_v1.json 2> gpurun_out/c3_bench_v1.err
B2Q_DECODE_V2=1 B2Q_DECODE2_KS=1 B2Q_DECODE2_GW=16 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c3_bench_v2.json 2> gpurun_out/c3_bench_v2.err
B2Q_DECODE2_XTMA=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c3_bench_hybrid_ldg.json 2> gpurun_out/c3_bench_hybrid_ldg.err
timeout 400 python tools/microbench.py gemv2 1 > gpurun_out/c3_gemv2_sweep.log 2>&1
for f in gpurun_out/c3_*tests.log gpurun_out/c3_san*.log; do echo "## $f: $(tail -1 $f | cut -c1-200)"; done
grep -E "FAIL|Error" gpurun_out/c3_san_moe.log gpurun_out/c3_san_midm.log | head
python - <<'PY'
import json
for f in ("hybrid","v1","v2","hybrid_ldg"):
    try:
        d = json.loads(open(f"gpurun_out/c3_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "decode tok/s", round(d["value"], 1), "frac", round(d["roofline"]["frac"],3), "prefill", round(d["prefill"]["tflops"],1))
    except Exception as e:
        print(f, "unreadable:", e)
PY
grep MIDM gpurun_out/c3_midm_bench.log | head -40
grep -E "^DECODE2" gpurun_out/c3_gemv2_sweep.log
