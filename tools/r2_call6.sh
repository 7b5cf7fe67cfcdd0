#!/bin/bash
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; ( "$@" ) > gpurun_out/c6_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/c6_$tag.log; echo "## $tag: $(grep -E 'ok$|rc=' gpurun_out/c6_$tag.log | tr '\n' ' ')"; }
run full_c36 timeout 120 python tools/san_midm_graph.py 4096 4096 16 36
run m64_c36 timeout 120 python tools/san_midm_graph.py 4096 4096 64 36
run m128_k14336 timeout 120 python tools/san_midm_graph.py 14336 4096 128 12
run n14336_c12 timeout 120 python tools/san_midm_graph.py 4096 14336 32 12
timeout 300 python tools/san_midm.py > gpurun_out/c6_san_midm.log 2>&1
timeout 300 python tools/san_moe.py > gpurun_out/c6_san_moe.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c6_default_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/c6_parity_default.json 2>/dev/null
timeout 300 python tools/microbench.py midm 16 64 128 > gpurun_out/c6_midm_bench.log 2>&1
timeout 1200 python bench.py --steps 30 --warmup 5 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
for f in gpurun_out/c6_san*.log gpurun_out/c6_*tests.log; do echo "## $f: $(tail -1 $f | cut -c1-220)"; done
grep MIDM gpurun_out/c6_midm_bench.log | head -44
tail -c 6000 gpurun_out/c6_bench.json
tail -5 gpurun_out/c6_bench.err
