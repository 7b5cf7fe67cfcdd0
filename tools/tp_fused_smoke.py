"""N-rank smoke of the fused row-parallel decode + all-reduce kernel (`b2q_decode_allreduce`, EXPERIMENTAL).

    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tools/tp_fused_smoke.py

Compares against (unfused matmul -> NCCL all-reduce) and the unsharded oracle, eagerly and under CUDA-graph replay,
for several token counts and layer shapes, then times fused vs P2P vs NCCL.  A watchdog dumps all stacks and exits
after 90 s so a protocol bug cannot hang the GPU box.
"""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(90, exit=True)
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
def P(*a):
    print(f"[r{rank} {time.time() % 1000:7.2f}]", *a, flush=True)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = torch.ones(8, device=dev); dist.all_reduce(t); torch.cuda.synchronize()
from gptqmodel_b200 import B200QuantLinear, tp
from helpers import random_layer, oracle_forward
far = tp.FusedDecodeAllReduce(dev, max_elems=8 * 8192)
par = tp.P2PAllReduce(dev, max_elems=8 * 8192)
P("buffers ok")
worst = 0.0
for (K, N, sym, gs) in ((2048, 1024, True, 128), (4096, 4096, True, 128), (14336, 4096, True, 128), (4096, 4096, False, 64)):
    if (K // world) % gs or (K // world) % 128:
        continue
    full = random_layer(K, N, seed=K + N, sym=sym, group_size=gs, device="cuda")
    torch.manual_seed(K + N)  # every rank must draw the SAME bias (rank 1's oracle was off by its own bias in the first run)
    full["bias"] = (torch.randn(N, device="cuda") * 0.1).to(torch.float16)
    sh = tp.shard_rows(full, rank, world)
    mod = B200QuantLinear.from_checkpoint_tensors(sh["qweight"], sh["qzeros"], sh["scales"], sh["g_idx"], 4, gs,
                                                  bias=sh.get("bias"), sym=sym, device=dev)
    for M in (1, 2, 5, 8):
        torch.manual_seed(100 + M)
        x = (torch.randn(M, K, device=dev) * 0.5).to(torch.float16)
        xs = x[:, rank * K // world:(rank + 1) * K // world].contiguous()
        ref = mod(xs); dist.all_reduce(ref)
        for rep in range(3):  # consecutive calls exercise both slots and the sequence counter
            got = mod.forward_allreduce(xs, far)
        torch.cuda.synchronize()
        orc = oracle_forward(full, x.cpu()).to(dev)
        e_ref = float((got.float() - ref.float()).abs().max())
        e_orc = float(((got.float() - orc.float()).abs() / (orc.float().abs() + orc.float().pow(2).mean().sqrt())).max())
        worst = max(worst, e_orc)
        P(f"K={K} N={N} sym={sym} g={gs} M={M}: |fused-nccl|max={e_ref:.4g} rel-vs-oracle={e_orc:.3g}")
assert worst < 2e-3, worst  # fp32 partials summed before ONE rounding; bias joins the fp32 sum
# CUDA-graph replay + timing on the Llama-3-8B o_proj / down_proj shard shapes
for (K, N) in ((4096, 4096), (14336, 4096)):
    full = random_layer(K, N, seed=7, device="cuda")
    sh = tp.shard_rows(full, rank, world)
    mod = B200QuantLinear.from_checkpoint_tensors(sh["qweight"], sh["qzeros"], sh["scales"], sh["g_idx"], 4, 128, device=dev)
    xs = (torch.randn(1, K // world, device=dev) * 0.5).to(torch.float16)
    fns = {"fused": lambda: mod.forward_allreduce(xs, far), "p2p": lambda: par(mod(xs)),
           "nccl": lambda: tp.all_reduce_sum_(mod(xs))}
    outs = {}
    for name, fn in fns.items():
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50): y = fn()
        g.replay(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): g.replay()
        e1.record(); torch.cuda.synchronize()
        outs[name] = y.clone()
        P(f"K={K} N={N} {name}: {e0.elapsed_time(e1) / 20 / 50 * 1e3:.2f} us per (shard matmul + all-reduce)")
        del g
    P("fused vs nccl after replay:", float((outs["fused"].float() - outs["nccl"].float()).abs().max()))
torch.cuda.synchronize(); dist.barrier(); P("done"); os._exit(0)
