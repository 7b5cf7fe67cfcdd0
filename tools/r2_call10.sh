#!/bin/bash
# default bench (exactly what the driver runs) + reference arm + ncu evidence of the DEFAULT kernels
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; echo "bench wall ${SECONDS}s rc=$?" >> gpurun_out/c10_bench.err
timeout 300 python bench.py --impl reference > gpurun_out/c10_bench_ref.json 2> gpurun_out/c10_bench_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode2_kernel -c 2 -o gpurun_out/r02_decode2 -f python tools/prof_one.py gemv 4096 28672 > gpurun_out/c10_ncu_decode2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2p_kernel -c 1 -o gpurun_out/r02_gemm2p_final -f python tools/prof_one.py gemm 4096 4096 2048 > gpurun_out/c10_ncu_gemm2p.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:midm_kernel -c 2 -o gpurun_out/r02_midm_final -f python tools/prof_one.py gemm 4096 4096 16 > gpurun_out/c10_ncu_midm.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-competitors > gpurun_out/c10_ncu_bench.log 2>&1
grep -E "^\[bench|wall" gpurun_out/c10_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c10_bench.json").read().strip().splitlines()[-1])
print("decode", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "prefill", round(d["prefill"]["tflops"],1), round(d["prefill"]["roofline"]["frac"],3), d["clocks"])
print("cpu_baseline", d.get("cpu_baseline"))
print(json.dumps(d.get("competitors"))[:3000])
r = json.loads(open("gpurun_out/c10_bench_ref.json").read().strip().splitlines()[-1])
print("ref arm", r["value"], r["cpu_baseline"]["cores"])
PY
