#!/bin/bash
# 2-GPU call: validate the fused row-parallel matmul + all-reduce launch, TP parity against the unsharded oracle, A/B bench
set -u
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/tp_fused_smoke.py > gpurun_out/tp2_fused_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/tp2_fused_smoke.log
timeout 500 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/tp2_tests.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 5 --no-extra > gpurun_out/tp2_bench_p2p.json 2> gpurun_out/tp2_bench_p2p.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 30 --warmup 5 --no-extra --fused-allreduce > gpurun_out/tp2_bench_fused.json 2> gpurun_out/tp2_bench_fused.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 30 --warmup 5 --no-extra --nccl-allreduce > gpurun_out/tp2_bench_nccl.json 2> gpurun_out/tp2_bench_nccl.err
tail -25 gpurun_out/tp2_fused_smoke.log | cut -c1-200
tail -3 gpurun_out/tp2_tests.log
python - <<'PY'
import json
for f in ("p2p","fused","nccl"):
    try:
        d = json.loads(open(f"gpurun_out/tp2_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "decode tok/s", round(d["value"], 1), "prefill", round(d["prefill"]["tflops"],1), d["config"]["parallelism"][:80])
    except Exception as e:
        print(f, "unreadable:", e)
PY
