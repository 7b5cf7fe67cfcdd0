#!/bin/bash
# round-2 fourth GPU call: small-batch tier fault in back-to-back launches (bisect), gemm2p one-block-per-warp dequant,
# bench with extras + competitors, TP worker sanity on 1 GPU is skipped
set -u
mkdir -p gpurun_out
timeout 120 python tools/san_midm_graph.py > gpurun_out/c4_chain_small.log 2>&1; echo "rc=$?" >> gpurun_out/c4_chain_small.log
timeout 120 python tools/san_midm_graph.py 4096 4096 16 36 > gpurun_out/c4_chain_full.log 2>&1; echo "rc=$?" >> gpurun_out/c4_chain_full.log
B2Q_DISABLE_PDL=1 timeout 120 python tools/san_midm_graph.py 4096 4096 16 36 > gpurun_out/c4_chain_full_nopdl.log 2>&1; echo "rc=$?" >> gpurun_out/c4_chain_full_nopdl.log
B2Q_MIDM_KS=1 timeout 120 python tools/san_midm_graph.py 4096 4096 16 36 > gpurun_out/c4_chain_full_ks1.log 2>&1; echo "rc=$?" >> gpurun_out/c4_chain_full_ks1.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tools/san_midm_graph.py 1024 512 16 16 > gpurun_out/c4_chain_memcheck.log 2>&1
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python tools/san_midm_graph.py 512 256 16 6 > gpurun_out/c4_chain_racecheck.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c4_default_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/c4_parity_default.json 2>/dev/null
timeout 200 python tools/microbench.py gemm 2048 > gpurun_out/c4_gemm_bench.log 2>&1
B2Q_DISABLE_PDL=1 timeout 300 python tools/microbench.py midm 16 64 128 > gpurun_out/c4_midm_bench_nopdl.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
for f in gpurun_out/c4_chain*.log gpurun_out/c4_*tests.log; do echo "## $f: $(tail -2 $f | tr '\n' ' ' | cut -c1-220)"; done
grep -E "GEMM M|cuBLAS" gpurun_out/c4_gemm_bench.log
grep MIDM gpurun_out/c4_midm_bench_nopdl.log | head -14
tail -c 3000 gpurun_out/c4_bench.json
tail -5 gpurun_out/c4_bench.err
