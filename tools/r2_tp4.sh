#!/bin/bash
# 4-GPU call: GPU-0 validation of the sibling-fused prefill launch, TP-4 parity test, bench at TP-4 with the Mixtral TP-4 extra
# (BASELINE configs[4]) and the overlapped prefill all-reduce A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tp4_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/tp4_parity.json 2>/dev/null
timeout 200 python tools/stress.py 60 > gpurun_out/tp4_stress.log 2>&1; echo "rc=$?" >> gpurun_out/tp4_stress.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-competitors --no-cpu-baseline > gpurun_out/tp4_bench_n1.json 2> gpurun_out/tp4_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/tp4_bench_default.json 2> gpurun_out/tp4_bench_default.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --steps 30 --warmup 5 --no-extra --overlap-chunks 1 > gpurun_out/tp4_bench_nooverlap.json 2> gpurun_out/tp4_bench_nooverlap.err
tail -3 gpurun_out/tp4_tests.log; tail -2 gpurun_out/tp4_stress.log; grep -h "FAIL\|TP_OK" gpurun_out/tp_gpu_worker_w4.log | head -5
python - <<'PY'
import json
for f in ("n1","default","nooverlap"):
    try:
        d = json.loads(open(f"gpurun_out/tp4_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "decode tok/s", round(d["value"], 1), "prefill", round(d["prefill"]["tflops"],1), "ms", round(d["prefill"]["ms_per_pass"],2), d["config"]["parallelism"][:70])
        if d.get("extra"): print(json.dumps(d["extra"])[:1500])
    except Exception as e:
        print(f, "unreadable:", e, open(f"gpurun_out/tp4_bench_{f}.err").read()[-800:])
PY
