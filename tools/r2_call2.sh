#!/bin/bash
# round-2 second GPU call: deep-ring small-batch tier, full suite, competitors incl. the reference's Marlin, decode v1/v2
# sweep + timelines, ncu evidence of the DEFAULT kernels
set -u
mkdir -p gpurun_out
timeout 300 python tools/san_midm.py > gpurun_out/c2_san_midm.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c2_default_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/c2_parity_default.json 2>/dev/null
timeout 300 python tools/microbench.py midm 16 64 128 > gpurun_out/c2_midm_bench.log 2>&1
timeout 900 python tools/competitors.py --json gpurun_out/c2_competitors.json > gpurun_out/c2_competitors.log 2>&1
timeout 500 python tools/microbench.py gemv2 1 > gpurun_out/c2_gemv2_sweep.log 2>&1
for v in 0 1; do for shape in "4096 14336" "14336 4096" "4096 4096"; do
  B2Q_DECODE_V2=$v timeout 120 python tools/trace_decode.py $shape >> gpurun_out/c2_trace_v$((v+1)).log 2>&1
done; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -c 2 -o gpurun_out/r02_decode -f python tools/prof_one.py gemv 4096 14336 > gpurun_out/c2_ncu_decode.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2p_kernel -c 1 -o gpurun_out/r02_gemm2p -f python tools/prof_one.py gemm 4096 4096 2048 > gpurun_out/c2_ncu_gemm2p.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:midm_kernel -c 2 -o gpurun_out/r02_midm -f python tools/prof_one.py gemm 4096 4096 64 > gpurun_out/c2_ncu_midm.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c2_ncu_bench.log 2>&1
for f in gpurun_out/c2_*tests.log gpurun_out/c2_san*.log gpurun_out/c2_competitors.log; do echo "## $f: $(tail -1 $f | cut -c1-200)"; done
grep MIDM gpurun_out/c2_midm_bench.log | head -40
grep -E "^DECODE2|^   ks" gpurun_out/c2_gemv2_sweep.log | head -60
grep -A3 "'stack'\|^b2q \|^marlin" gpurun_out/c2_competitors.log | tail -12
