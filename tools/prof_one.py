"""Run a few launches of one tier on one layer (target for ncu).  usage: prof_one.py gemv|gemm K N [M]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gptqmodel_b200 import B200QuantLinear  # noqa: E402
from helpers import random_layer  # noqa: E402

what, K, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
M = int(sys.argv[4]) if len(sys.argv) > 4 else (1 if what == "gemv" else 2048)
mods = []
for c in range(6):
    L = random_layer(K, N, seed=c, device="cuda")
    mods.append(B200QuantLinear.from_checkpoint_tensors(L["qweight"], L["qzeros"], L["scales"], L["g_idx"], 4, 128,
                                                        device="cuda"))
x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.float16)
for m in mods:
    y = m(x)
torch.cuda.synchronize()
print("done", float(y.float().abs().mean()))
