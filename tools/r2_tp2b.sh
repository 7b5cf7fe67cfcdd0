#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/tp2b_tests.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/tp_fused_smoke.py > gpurun_out/tp2b_fused_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/tp2b_fused_smoke.log
for f in p2p fused; do
  flag=""; [ $f = fused ] && flag="--fused-allreduce"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 400)) bench.py --gpus 2 --steps 30 --warmup 5 --no-extra $flag > gpurun_out/tp2b_bench_$f.json 2> gpurun_out/tp2b_bench_$f.err
done
tail -3 gpurun_out/tp2b_tests.log; grep -h "TP_OK\|FAIL" gpurun_out/tp_gpu_worker_w2.log | head -4
grep -E "us per|done|rc=|Error|assert" gpurun_out/tp2b_fused_smoke.log | tail -12 | cut -c1-200
python - <<'PY'
import json
for f in ("p2p","fused"):
    try:
        d = json.loads(open(f"gpurun_out/tp2b_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "decode tok/s", round(d["value"], 1), "prefill", round(d["prefill"]["tflops"],1))
    except Exception as e:
        print(f, "unreadable:", e, open(f"gpurun_out/tp2b_bench_{f}.err").read()[-600:])
PY
