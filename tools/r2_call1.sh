#!/bin/bash
# round-2 first GPU call: new small-batch tier first contact, full default suite (no -x), competitors, then the
# never-run experimental kernels of round 1
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.log 2>&1
timeout 300 python tools/san_midm.py > gpurun_out/c1_san_midm.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python tools/san_midm.py > gpurun_out/c1_san_midm_memcheck.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c1_default_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/c1_parity_default.json 2>/dev/null
timeout 400 python tools/microbench.py midm > gpurun_out/c1_midm_bench.log 2>&1
timeout 900 python tools/competitors.py --json gpurun_out/c1_competitors.json > gpurun_out/c1_competitors.log 2>&1
( export B2Q_DECODE_V2=1 B2Q_GEMM2_STREAMK=1; timeout 300 python tools/san_one.py > gpurun_out/c1_san_plain.log 2>&1 )
B2Q_DECODE_V2=1 timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > gpurun_out/c1_v2_tests.log 2>&1
B2Q_GEMM2_STREAMK=1 timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "full_size or cases or 70b or act_order or oracle_fp16" > gpurun_out/c1_sk_tests.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c1_bench_v1.json 2> gpurun_out/c1_bench_v1.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --decode-v2 --gemm-streamk > gpurun_out/c1_bench_v2sk.json 2> gpurun_out/c1_bench_v2sk.err
for f in gpurun_out/c1_*tests.log gpurun_out/c1_san*.log gpurun_out/c1_competitors.log; do echo "## $f: $(tail -1 $f | cut -c1-200)"; done
tail -c 600 gpurun_out/c1_bench_v1.json; echo; tail -c 600 gpurun_out/c1_bench_v2sk.json
grep MIDM gpurun_out/c1_midm_bench.log | head -60
