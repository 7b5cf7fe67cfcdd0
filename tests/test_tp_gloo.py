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
        # act-order row shards are refused loudly
        ao = make_layer(256, 128, group_size=64, desc_act=True, seed=3)
        try:
            tp.shard_rows(ao, rank, world)
            ok = False
        except NotImplementedError:
            pass
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
