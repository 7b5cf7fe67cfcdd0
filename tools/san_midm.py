"""Small-shape exercise of the small-batch tier (b2q_midm.cu): quick parity check and compute-sanitizer target.

    python tools/san_midm.py
    compute-sanitizer --tool memcheck python tools/san_midm.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gptqmodel_b200 as g  # noqa: E402
from gptqmodel_b200 import B200QuantLinear  # noqa: E402
from helpers import assert_close_rel, make_layer, oracle_forward  # noqa: E402


def mod(L):
    return B200QuantLinear.from_checkpoint_tensors(L["qweight"], L["qzeros"], L["scales"], L["g_idx"], L["bits"],
                                                   L["group_size"], bias=L["bias"], desc_act=L["desc_act"], sym=L["sym"])


fails = 0
layers = [make_layer(512, 256, bits=4, group_size=128, sym=True, seed=1),
          make_layer(1024, 160, bits=4, group_size=32, sym=False, bias=True, seed=2),
          make_layer(512, 256, bits=8, group_size=64, sym=False, seed=3),
          make_layer(1024, 128, bits=4, group_size=64, sym=True, desc_act=True, seed=4)]
for L in layers:
    m = mod(L)
    for ks in (0, 1, 2, 4, 8):
        if ks:
            os.environ["B2Q_MIDM_KS"] = str(ks)
        else:
            os.environ.pop("B2Q_MIDM_KS", None)
        g.lib.b2q_debug_reload_env()
        for M in (9, 16, 33, 100, 128):
            x = (torch.randn(M, L["K"], generator=torch.Generator().manual_seed(M)) * 0.5).to(torch.float16)
            what = f"midm bits={L['bits']} K={L['K']} N={L['N']} g={L['group_size']} act={L['desc_act']} ks={ks} M={M}"
            try:
                y = m(x.cuda())
                torch.cuda.synchronize()
                assert_close_rel(y, oracle_forward(L, x), 1e-3, what)
                print("ok", what, flush=True)
            except AssertionError as e:
                fails += 1
                print("FAIL", what, str(e)[:200], flush=True)
os.environ.pop("B2Q_MIDM_KS", None)
print("all ok" if fails == 0 else f"{fails} FAILED")
sys.exit(1 if fails else 0)
