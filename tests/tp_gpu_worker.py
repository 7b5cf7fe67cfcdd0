"""torchrun worker of tests/test_tp_gpu.py: tensor-parallel QuantLinear stack on real GPUs against the UNSHARDED oracle.

One rank per GPU (NCCL).  A Llama-style MLP pair — up (column-parallel) -> down (row-parallel, ONE all-reduce) — is sharded
with gptqmodel_b200.tp, run through the B200 kernels and compared with the oracle evaluated on the unsharded layer, for the
decode tier (M = 1, 5), the small-batch tier (M = 40) and the prefill tier (M = 300), with all three reductions: NCCL,
the one-shot peer-memory kernel (b2q_allreduce) and the fused matmul + all-reduce launch (b2q_decode_allreduce, M <= 8).
A watchdog aborts after 150 s so a protocol bug cannot hang the box.
"""
import faulthandler
import os
import sys

faulthandler.dump_traceback_later(150, exit=True)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from gptqmodel_b200 import B200QuantLinear, tp  # noqa: E402
from helpers import assert_close_rel, make_layer, oracle_forward  # noqa: E402


def mod(L, sym, gs):
    return B200QuantLinear.from_checkpoint_tensors(L["qweight"], L["qzeros"], L["scales"], L["g_idx"], 4, gs,
                                                   bias=L.get("bias"), sym=sym, device=dev)


far = par = None
try:
    par = tp.P2PAllReduce(dev, max_elems=8 * 2048)
    far = tp.FusedDecodeAllReduce(dev, max_elems=8 * 2048)
except Exception as e:  # noqa: BLE001  (no symmetric memory on this box: NCCL only)
    if rank == 0:
        print("symmetric memory unavailable:", type(e).__name__, e, flush=True)
worst = {}
for sym, gs, bias in ((True, 128, False), (False, 64, True)):
    hidden, inter = 1024, 2048 * world // 2 if world > 2 else 2048
    up = make_layer(hidden, inter, group_size=gs, sym=sym, bias=bias, seed=31)
    down = make_layer(inter, hidden, group_size=gs, sym=sym, bias=bias, seed=32)
    mu = mod(tp.shard_columns(up, rank, world), sym, gs)
    md = mod(tp.shard_rows(down, rank, world), sym, gs)
    for M in (1, 5, 40, 300):
        x = (torch.randn(M, hidden, generator=torch.Generator().manual_seed(M)) * 0.5).to(torch.float16)
        h_ref = oracle_forward(up, x)
        y_ref = oracle_forward(down, h_ref)
        h = mu(x.to(dev))                                   # [M, inter / world]: this rank's K-slice of `down`
        assert_close_rel(h, h_ref[:, rank * inter // world:(rank + 1) * inter // world], 1e-3, f"column shard M={M}")
        reducers = {"nccl": tp.RowParallelLinear(md)}
        if par is not None:
            reducers["p2p"] = tp.RowParallelLinear(md, reduce=par)
        if far is not None and M <= 8:
            reducers["fused"] = tp.RowParallelLinear(md, reduce=far)
        for name, layer in reducers.items():
            for rep in range(2):                            # twice: both slots / the sequence counters of our kernels
                y = layer(h_ref[:, rank * inter // world:(rank + 1) * inter // world].contiguous().to(dev))
            torch.cuda.synchronize()
            # NCCL / p2p: every rank's partial sum is rounded to 16 bits and the reduction adds `world` of them (NCCL's ring
            # rounds the running sum at every hop): ~sqrt(2 * world) roundings of 2^-11 instead of one -> 1e-3 * (1 +
            # sqrt(world) / 2).  The fused launch sums fp32 partials and rounds once: 2e-3 covers its chained oracle.
            rel = 2e-3 if name == "fused" else 1e-3 * (1.0 + world ** 0.5 / 2)
            try:
                assert_close_rel(y, y_ref, rel, f"{name} sym={sym} M={M} world={world}")
            except AssertionError as e:
                print(f"[r{rank}] FAIL {e}", flush=True)
                raise
            o, r = y.float().cpu(), y_ref.float()
            ratio = float(((o - r).abs() / (rel * r.abs() + rel * r.pow(2).mean().sqrt())).max())
            worst[name] = max(worst.get(name, 0.0), ratio)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("TP_OK world", world, {k: round(v, 3) for k, v in worst.items()}, flush=True)
os._exit(0)
