#!/bin/bash
# Round-2 checklist for the three code paths that were written in round 1 AFTER the GPU budget ran out (they compile
# for sm_100a, their host logic and index arithmetic are covered by CPU tests, but they have never run on a GPU).
# Run each block as ONE gpurun call; every command is wrapped in `timeout` so a protocol bug cannot hang the box.
#
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/validate_experimental.sh all'        (first call: ~35 GPU-min)
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/validate_experimental.sh sanitize'   (if `all` shows failures)
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/validate_experimental.sh decode'   (~20 GPU-min)
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/validate_experimental.sh gemm'
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/validate_experimental.sh midm'
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/validate_experimental.sh profile'
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/validate_experimental.sh tp 2'
set -u
mkdir -p gpurun_out
summary() {  # one line per log: name + last line
  for f in "$@"; do [ -f "$f" ] && echo "## $f: $(tail -1 "$f" | cut -c1-220)"; done
}
case "${1:-decode}" in
  all)
    # everything a first GPU call should answer, bounded: small-shape parity of all experimental kernels, then the GPU
    # suite under each switch, then A/B bench lines.  ~35 GPU-min.  If san_plain fails, run `sanitize` next.
    ( export B2Q_DECODE_V2=1 B2Q_GEMM2_STREAMK=1 B2Q_GEMM_SPLITK=1; timeout 300 python tools/san_one.py > gpurun_out/san_plain.log 2>&1 )
    B2Q_DECODE_V2=1 timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/v2_tests.log 2>&1
    B2Q_GEMM2_STREAMK=1 timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/sk_tests.log 2>&1
    B2Q_GEMM_SPLITK=1 timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gsk_tests.log 2>&1
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --decode-v2 --gemm-streamk > gpurun_out/bench_v2sk.json 2> gpurun_out/bench_v2sk.err
    for sk in 0 1; do B2Q_GEMM_SPLITK=$sk timeout 200 python tools/microbench.py gemm 17 64 128 > gpurun_out/gsk_bench_$sk.log 2>&1; done
    summary gpurun_out/san_plain.log gpurun_out/v2_tests.log gpurun_out/sk_tests.log gpurun_out/gsk_tests.log
    python - <<'PY'
import json
for f in ("gpurun_out/bench_v1.json", "gpurun_out/bench_v2sk.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "decode tok/s", round(d["value"], 1), "roofline", d["roofline"]["frac"], "prefill", d.get("prefill"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
    grep -h "GEMM M=" gpurun_out/gsk_bench_0.log gpurun_out/gsk_bench_1.log | head -40
    ;;
  decode)
    # 1. parity: the whole GPU suite with every decode launch routed through b2q_decode2.cu
    B2Q_DECODE_V2=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/v2_tests.log
    # 1b. the LDG staging path (B2Q_DECODE2_XTMA=0) and forced warp groups
    B2Q_DECODE_V2=1 B2Q_DECODE2_XTMA=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decode or cases or sibling" 2>&1 | tail -5 | tee -a gpurun_out/v2_tests.log
    for gw in 4 8; do
      B2Q_DECODE_V2=1 B2Q_DECODE2_GW=$gw timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decode or cases or sibling" 2>&1 | tail -3 | tee -a gpurun_out/v2_tests.log
    done
    B2Q_DECODE_V2=1 B2Q_DECODE2_FASTSYNC=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "decode or cases or sibling or llama" 2>&1 | tail -3 | tee -a gpurun_out/v2_tests.log
    # 2. A/B: v1 vs v2 on the Llama-3-8B stack (same box, back to back)
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --decode-v2 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
    B2Q_DECODE2_FASTSYNC=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --decode-v2 > gpurun_out/bench_v2_fastsync.json 2> gpurun_out/bench_v2_fastsync.err
    B2Q_DECODE2_XTMA=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --decode-v2 > gpurun_out/bench_v2_ldg.json 2> gpurun_out/bench_v2_ldg.err
    for f in v1 v2 v2_fastsync v2_ldg; do echo "== $f"; tail -c 700 gpurun_out/bench_$f.json | head -c 700; echo; done
    # 2b. per-shape sweep of v2's launch parameters against its planner's choice (cost-model calibration)
    timeout 400 python tools/microbench.py gemv2 1 2>&1 | tail -50 | tee gpurun_out/v2_sweep.log
    # 3. phase timelines of the two kernels on the widest layers (ns; min / median / max over CTAs)
    for v in 0 1; do
      for shape in "4096 14336" "14336 4096"; do
        B2Q_DECODE_V2=$v timeout 120 python tools/trace_decode.py $shape 2>&1 | tail -24 | tee -a gpurun_out/trace_v$((v+1)).log
      done
    done
    ;;
  gemm)
    B2Q_GEMM2_STREAMK=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/sk_tests.log
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp.json 2> gpurun_out/bench_dp.err
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-streamk > gpurun_out/bench_sk.json 2> gpurun_out/bench_sk.err
    python - <<'PY'
import json
for f in ("gpurun_out/bench_dp.json", "gpurun_out/bench_sk.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get("prefill"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
    ;;
  midm)
    # cluster split-K of the single-CTA tcgen05 tier (M <= 128 and all 8-bit shapes)
    B2Q_GEMM_SPLITK=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gsk_tests.log
    for sk in 0 1; do
      B2Q_GEMM_SPLITK=$sk timeout 200 python tools/microbench.py gemm 17 32 64 128 2>&1 | tail -24 | tee gpurun_out/gsk_bench_$sk.log
    done
    ;;
  sanitize)
    # first contact with the hardware: small shapes under compute-sanitizer (memcheck finds the out-of-bounds smem /
    # global accesses an index slip would cause, synccheck the barrier misuse, racecheck shared-memory hazards)
    export B2Q_DECODE_V2=1 B2Q_GEMM2_STREAMK=1 B2Q_GEMM_SPLITK=1
    timeout 300 python tools/san_one.py 2>&1 | tail -15 | tee gpurun_out/san_plain.log
    for tool in memcheck synccheck racecheck; do
      timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/san_one.py 2>&1 | tail -40 | tee gpurun_out/san_$tool.log
    done
    ;;
  profile)
    # ncu captures of the new kernels (one GPU; numbers printed under ncu are never bench values).  Read them in the
    # container with `ncu -i <rep> --page raw --csv` and summarise into profiles/r02_*.txt.
    B2Q_DECODE_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode2_kernel -c 3 \
        -o gpurun_out/r02_decode2 -f python tools/prof_one.py gemv 4096 14336 > gpurun_out/ncu_decode2.log 2>&1
    B2Q_GEMM2_STREAMK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2s_kernel -c 2 \
        -o gpurun_out/r02_gemm2s -f python tools/prof_one.py gemm 4096 4096 2048 > gpurun_out/ncu_gemm2s.log 2>&1
    B2Q_DECODE_V2=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --decode-v2 > gpurun_out/ncu_bench.log 2>&1
    ls -la gpurun_out | tail -8
    ;;
  tp)
    N=${2:-2}
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
        tools/tp_fused_smoke.py 2>&1 | tail -60 | tee gpurun_out/tp_fused_smoke.log
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29518 \
        bench.py --gpus "$N" --steps 30 --warmup 5 --no-cpu-baseline --decode-v2 --fused-allreduce 2>&1 | tail -3 | tee gpurun_out/bench_tp_fused.log
    ;;
esac
