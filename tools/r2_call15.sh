#!/bin/bash
# final validation of the round: smoke, full GPU suite, sanitizer on the newest launch paths, stress, kernel-filtered launch
# list of the timed bench, default bench + reference arm
set -u
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/c15_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c15_smoke.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c15_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/c15_parity.json 2>/dev/null
timeout 500 compute-sanitizer --tool memcheck --print-limit 10 python tools/san_one.py > gpurun_out/c15_san_one_memcheck.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python tools/san_moe.py > gpurun_out/c15_san_moe_memcheck.log 2>&1
timeout 300 python tools/stress.py 150 > gpurun_out/c15_stress.log 2>&1; echo "rc=$?" >> gpurun_out/c15_stress.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode|gemm2p|midm|allreduce|gemv_kernel|gemm_kernel" -c 520 --csv --log-file gpurun_out/r02_launches_c15.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-competitors > gpurun_out/c15_ncu_bench.log 2>&1
timeout 900 python bench.py > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
timeout 300 python bench.py --impl reference > gpurun_out/c15_bench_ref.json 2> gpurun_out/c15_bench_ref.err
for f in gpurun_out/c15_smoke.log gpurun_out/c15_tests.log gpurun_out/c15_san_one_memcheck.log gpurun_out/c15_stress.log; do echo "## $f: $(tail -2 $f | tr '\n' ' ' | cut -c1-200)"; done
tail -3 gpurun_out/c15_san_moe_memcheck.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c15_bench.json").read().strip().splitlines()[-1])
print("decode", round(d["value"],1), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "prefill", round(d["prefill"]["tflops"],1), round(d["prefill"]["roofline"]["frac"],3), d["clocks"])
r = json.loads(open("gpurun_out/c15_bench_ref.json").read().strip().splitlines()[-1])
print("ref arm", r["value"], r["cpu_baseline"]["cores"])
PY
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c15_bench.json").read().strip().splitlines()[-1])
print("mixtral", json.dumps(d["extra"]["mixtral_8x7b_tp4"])[:500])
PY
grep -n "FAILED\|Error" gpurun_out/c15_tests.log | head
