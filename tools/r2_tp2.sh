#!/bin/bash
# 2-GPU call: 4-issuer small-batch tier validation (GPU 0), then the fused row-parallel matmul + all-reduce launch, TP parity
# against the unsharded oracle (torchrun), A/B bench at TP-2
set -u
mkdir -p gpurun_out
timeout 120 python tools/san_midm_graph.py 4096 4096 16 36 > gpurun_out/tp2_chain.log 2>&1; echo "rc=$?" >> gpurun_out/tp2_chain.log
timeout 300 python tools/san_midm.py > gpurun_out/tp2_san_midm.log 2>&1
timeout 300 python tools/san_moe.py > gpurun_out/tp2_san_moe.log 2>&1
timeout 300 python tools/stress.py 100 > gpurun_out/tp2_stress.log 2>&1; echo "rc=$?" >> gpurun_out/tp2_stress.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/tp2_tests.log 2>&1
cp gpurun_out/parity.json gpurun_out/tp2_parity.json 2>/dev/null
timeout 200 python tools/microbench.py midm 16 128 > gpurun_out/tp2_midm_bench.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/tp_fused_smoke.py > gpurun_out/tp2_fused_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/tp2_fused_smoke.log
for f in p2p fused nccl; do
  flag=""; [ $f = fused ] && flag="--fused-allreduce"; [ $f = nccl ] && flag="--nccl-allreduce"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 400)) bench.py --gpus 2 --steps 30 --warmup 5 --no-extra $flag > gpurun_out/tp2_bench_$f.json 2> gpurun_out/tp2_bench_$f.err
done
for f in gpurun_out/tp2_chain.log gpurun_out/tp2_san_midm.log gpurun_out/tp2_san_moe.log gpurun_out/tp2_stress.log gpurun_out/tp2_tests.log; do echo "## $f: $(tail -2 $f | tr '\n' ' ' | cut -c1-220)"; done
grep MIDM gpurun_out/tp2_midm_bench.log | head -12
tail -22 gpurun_out/tp2_fused_smoke.log | cut -c1-200
python - <<'PY'
import json
for f in ("p2p","fused","nccl"):
    try:
        d = json.loads(open(f"gpurun_out/tp2_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, "decode tok/s", round(d["value"], 1), "prefill", round(d["prefill"]["tflops"],1), d["config"]["parallelism"][:90])
    except Exception as e:
        print(f, "unreadable:", e, open(f"gpurun_out/tp2_bench_{f}.err").read()[-600:])
PY
