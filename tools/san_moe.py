"""Small MoE block through the grouped kernels (compute-sanitizer target / quick parity check)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gptqmodel_b200 import B200QuantLinear, moe  # noqa: E402
from helpers import assert_close_rel, make_layer, oracle_forward  # noqa: E402

E, K, I, top_k = 4, 256, 384, 2
fails = 0
for sym, gs in ((False, 64), (True, 128)):
    layers = [(make_layer(K, I, group_size=gs, sym=sym, seed=200 + 3 * e), make_layer(K, I, group_size=gs, sym=sym, seed=201 + 3 * e),
               make_layer(I, K, group_size=gs, sym=sym, seed=202 + 3 * e)) for e in range(E)]
    mk = lambda L: B200QuantLinear.from_checkpoint_tensors(L["qweight"], L["qzeros"], L["scales"], L["g_idx"], 4, gs, sym=sym)  # noqa: E731
    blk = moe.MoEExperts([mk(l[0]) for l in layers], [mk(l[1]) for l in layers], [mk(l[2]) for l in layers], grouped=True)
    for T in (1, 5, 20, 70):
        gen = torch.Generator().manual_seed(T)
        x = (torch.randn(T, K, generator=gen) * 0.5).to(torch.float16)
        ids, w = moe.route_topk(torch.randn(T, E, generator=gen), top_k)
        ref = torch.zeros(T, K)
        for t in range(T):
            for j in range(top_k):
                l1, l3, l2 = layers[int(ids[t, j])]
                xt = x[t:t + 1]
                h = (F.silu(oracle_forward(l1, xt).float()) * oracle_forward(l3, xt).float()).to(torch.float16)
                ref[t] += float(w[t, j]) * oracle_forward(l2, h)[0].float()
        try:
            got = blk(x.cuda(), ids.cuda(), w.cuda())
            torch.cuda.synchronize()
            assert_close_rel(got, ref, 4e-3, f"moe sym={sym} g={gs} T={T}")
            print("ok moe", sym, gs, T, flush=True)
        except AssertionError as e:
            fails += 1
            print("FAIL moe", sym, gs, T, str(e)[:300], flush=True)
print("all ok" if fails == 0 else f"{fails} FAILED")
sys.exit(1 if fails else 0)
