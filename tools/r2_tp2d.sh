#!/bin/bash
# 2-GPU call after the tp.shard_rows / FusedDecodeAllReduce / act-order staging changes: TP parity (torchrun inside the test),
# fused row-parallel smoke, TP-2 bench (default reduction)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity_formats.py -m gpu -q -p no:cacheprovider -k "moe" > gpurun_out/tp2d_t1.log 2>&1
timeout 600 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/tp2d_tp.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/tp_fused_smoke.py > gpurun_out/tp2d_fused_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/tp2d_fused_smoke.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --no-extra --no-competitors > gpurun_out/tp2d_bench.json 2> gpurun_out/tp2d_bench.err
for f in gpurun_out/tp2d_t1.log gpurun_out/tp2d_tp.log gpurun_out/tp2d_fused_smoke.log; do echo "## $f: $(tail -3 $f | tr '\n' ' ' | cut -c1-300)"; done
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/tp2d_bench.json").read().strip().splitlines()[-1])
    print("TP-2 decode tok/s", round(d["value"], 1), "prefill", round(d["prefill"]["tflops"],1), d["config"]["parallelism"][:90])
except Exception as e:
    print("unreadable:", e, open("gpurun_out/tp2d_bench.err").read()[-800:])
PY
