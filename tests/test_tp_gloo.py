"""CPU, world_size=2 over gloo: column/row sharding + the row-parallel all-reduce reproduce the unsharded layer.

The oracle stands in for the kernel on each rank (there is no CPU product path); what is under test is the host
logic of gptqmodel_b200/tp.py: slicing of qweight / qzeros / scales / g_idx, bias-on-rank-0, and the collective.
"""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from gptqmodel_b200 import tp
        from helpers import make_layer

        torch.manual_seed(0)
        ok = True
        for bits, gs, sym, bias in ((4, 128, True, False), (4, 64, False, True), (8, 32, False, True),
                                    (4, -1, True, False)):
            up = make_layer(256, 512, bits=bits, group_size=gs, sym=sym, bias=bias, seed=11)    # column-parallel
            down = make_layer(512, 256, bits=bits, group_size=gs, sym=sym, bias=bias, seed=12)  # row-parallel
            x = (torch.randn(3, 256, generator=torch.Generator().manual_seed(5)) * 0.5).to(torch.float16)

            def fwd(L, inp):
                return oracle.forward(inp, L["qweight"], L["qzeros"], L["scales"], L["g_idx"], L["bits"],
                                      bias=L["bias"])

            full = fwd(down, fwd(up, x))
            us, ds = tp.shard_columns(up, rank, world), tp.shard_rows(down, rank, world)
            assert us["qweight"].shape == (256 * bits // 32, 512 // world)
            assert ds["qweight"].shape == (512 // world * bits // 32, 256)
            h = fwd(us, x)                      # [3, 512/world] — exactly the K-slice the row shard consumes
            part = fwd(ds, h).float()           # partial sums (bias only on rank 0)
            tp.all_reduce_sum_(part)
            err = (part - full.float()).abs().max().item()
            ok = ok and err < 2e-2 * full.float().abs().max().item()
            # column shards concatenate to the unsharded output
            outs = [torch.zeros_like(h) for _ in range(world)]
            dist.all_gather(outs, h)
            ok = ok and torch.equal(torch.cat(outs, dim=1), fwd(up, x))
        # act-order row shards keep the FULL scale / zero tables and their slice of g_idx (utils/marlin.py:300-305) ...
        ao = make_layer(256, 128, group_size=64, desc_act=True, seed=3)
        x = (torch.randn(3, 256, generator=torch.Generator().manual_seed(6)) * 0.5).to(torch.float16)
        rs = tp.shard_rows(ao, rank, world)
        ok = ok and rs["replicated_tables"] and rs["scales"].shape == ao["scales"].shape
        ok = ok and torch.equal(rs["g_idx"], ao["g_idx"][rank * 256 // world:(rank + 1) * 256 // world])
        part = oracle.forward(x[:, rank * 256 // world:(rank + 1) * 256 // world].contiguous(), rs["qweight"], rs["qzeros"],
                              rs["scales"], rs["g_idx"], rs["bits"], bias=rs["bias"]).float()
        tp.all_reduce_sum_(part)
        full = oracle.forward(x, ao["qweight"], ao["qzeros"], ao["scales"], ao["g_idx"], ao["bits"], bias=ao["bias"])
        ok = ok and (part - full.float()).abs().max().item() < 2e-2 * full.float().abs().max().item()
        # ... or are served column-parallel between two all-gathers (same interface: K/P in, full N out)
        cs = tp.shard_columns(ao, rank, world)

        class Inner(torch.nn.Module):
            def forward(self, inp):
                return oracle.forward(inp, cs["qweight"], cs["qzeros"], cs["scales"], cs["g_idx"], cs["bits"], bias=cs["bias"])

        wrapped = tp.GatheredColumnParallelLinear(Inner())
        got = wrapped(x[:, rank * 256 // world:(rank + 1) * 256 // world].contiguous())
        full = oracle.forward(x, ao["qweight"], ao["qzeros"], ao["scales"], ao["g_idx"], ao["bits"], bias=ao["bias"])
        ok = ok and torch.equal(got, full)
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_tp2_sharding_and_allreduce_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def _moe_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        import torch.nn.functional as F
        from gptqmodel_b200 import moe, tp
        from helpers import make_layer

        E, K, I, T, top_k = 4, 128, 256, 9, 2     # Mixtral-style block in miniature: g64 asymmetric experts
        layers = [(make_layer(K, I, group_size=64, sym=False, seed=100 + 3 * e),
                   make_layer(K, I, group_size=64, sym=False, seed=101 + 3 * e),
                   make_layer(I, K, group_size=64, sym=False, seed=102 + 3 * e)) for e in range(E)]

        def dense(L):  # the oracle stands in for the kernel: [m, K] -> [m, N]
            return lambda inp: oracle.forward(inp, L["qweight"], L["qzeros"], L["scales"], L["g_idx"], L["bits"],
                                              bias=L["bias"])

        gen = torch.Generator().manual_seed(4)
        x = (torch.randn(T, K, generator=gen) * 0.5).to(torch.float16)
        ids, w = moe.route_topk(torch.randn(T, E, generator=gen), top_k)
        assert ids.shape == (T, top_k) and torch.allclose(w.sum(-1), torch.ones(T))
        # straightforward per-token reference on the unsharded experts
        ref = torch.zeros(T, K)
        for t in range(T):
            for j in range(top_k):
                w1, w3, w2 = (dense(L) for L in layers[int(ids[t, j])])
                xt = x[t:t + 1]
                ref[t] += float(w[t, j]) * w2(F.silu(w1(xt)) * w3(xt))[0].float()
        shards = [tp.shard_moe_expert(*layers[e], rank, world) for e in range(E)]
        assert shards[0][0]["qweight"].shape[1] == I // world and shards[0][2]["qweight"].shape[0] == I // world * 4 // 32
        blk = moe.MoEExperts([dense(s[0]) for s in shards], [dense(s[1]) for s in shards], [dense(s[2]) for s in shards])
        got = blk(x, ids, w)                        # ends in ONE all-reduce over the TP group
        err = (got.float() - ref).abs().max().item()
        ok = err < 2e-2 * ref.abs().max().item()
        # every rank holds the same result; an expert nobody routed to is skipped
        outs = [torch.zeros_like(got) for _ in range(world)]
        dist.all_gather(outs, got)
        ok = ok and torch.equal(outs[0], outs[1])
        only0 = torch.zeros(T, top_k, dtype=torch.long)
        only0[:, 1] = 1
        got2 = blk(x, only0, torch.full((T, top_k), 0.5))
        ok = ok and torch.isfinite(got2.float()).all().item()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_moe_block_tp2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_moe_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
