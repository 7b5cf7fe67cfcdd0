#!/bin/bash
# 8-GPU call: TP-8 parity against the unsharded oracle, SCALE-style bench (default flags) with the 70B TP-8 extra, fused A/B
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tp_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/tp8_tests.log 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/tp8_bench_default.json 2> gpurun_out/tp8_bench_default.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 30 --warmup 5 --no-extra --fused-allreduce > gpurun_out/tp8_bench_fused.json 2> gpurun_out/tp8_bench_fused.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 30 --warmup 5 --no-extra --nccl-allreduce > gpurun_out/tp8_bench_nccl.json 2> gpurun_out/tp8_bench_nccl.err
tail -3 gpurun_out/tp8_tests.log
python - <<'PY'
import json
for f in ("default","fused","nccl"):
    try:
        d = json.loads(open(f"gpurun_out/tp8_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "decode tok/s", round(d["value"], 1), "prefill", round(d["prefill"]["tflops"],1), "ms", round(d["prefill"]["ms_per_pass"],2), d["config"]["parallelism"][:90])
        if d.get("extra"): print(json.dumps(d["extra"])[:1500])
    except Exception as e:
        print(f, "unreadable:", e, open(f"gpurun_out/tp8_bench_{f}.err").read()[-800:])
PY
grep "^\[bench" gpurun_out/tp8_bench_default.err
