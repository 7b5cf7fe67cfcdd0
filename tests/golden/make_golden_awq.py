#!/usr/bin/env python
"""Golden fixtures for the AWQ (FORMAT.GEMM) front-end, produced by the UNMODIFIED reference.

Run in the authoring container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_awq.py

Output (committed): tests/golden/awq_cases.npz — for every case the AWQ-layout tensors (qweight int32 [K, N/8] with
the interleaved nibble order, qzeros int32 [G, N/8], scales fp16 [G, N], optional bias), the dequantised weight from
the reference's `dequantize_gemm` (gptqmodel/quantization/awq/utils/packing_utils.py:106-121) and the fp16 / bf16
forward outputs of the reference's `AwqTorchLinear` (gptqmodel/nn_modules/qlinear/torch_awq.py:157-197).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _StubFinder, _shell, REF_ROOT  # noqa: E402

# (name, group_size, K, N, M, bias)
CASES = [
    ("awq_g128", 128, 256, 128, 5, False),
    ("awq_g32_bias", 32, 128, 64, 3, True),
    ("awq_g64", 64, 256, 64, 1, True),
    ("awq_gK", -1, 128, 64, 2, False),
]


def main():
    sys.meta_path.append(_StubFinder())
    g = _shell("gptqmodel", REF_ROOT + "/gptqmodel")
    g.DEBUG_ON = False
    _shell("gptqmodel.models", REF_ROOT + "/gptqmodel/models")
    from gptqmodel.nn_modules.qlinear.torch_awq import AwqTorchLinear
    from gptqmodel.quantization.awq.utils.packing_utils import dequantize_gemm

    blobs, meta = {}, {}
    for i, (name, gs, K, N, M, bias) in enumerate(CASES):
        gen = torch.Generator().manual_seed(7000 + i)
        eff = gs if gs > 0 else K
        G = K // eff
        qweight = torch.randint(-(2 ** 31), 2 ** 31 - 1, (K, N // 8), dtype=torch.int32, generator=gen)
        qzeros = torch.randint(-(2 ** 31), 2 ** 31 - 1, (G, N // 8), dtype=torch.int32, generator=gen)
        scales = (torch.rand(G, N, generator=gen) * 0.02 + 0.005).to(torch.float16)
        b = (torch.randn(N, generator=gen) * 0.1).to(torch.float16) if bias else None
        mod = AwqTorchLinear(bits=4, group_size=gs, sym=False, desc_act=False, in_features=K, out_features=N, bias=bias,
                             register_buffers=True)
        with torch.no_grad():
            mod.qweight.copy_(qweight)
            mod.qzeros.copy_(qzeros)
            mod.scales.copy_(scales)
            if bias:
                mod.bias.copy_(b)
        mod.eval()
        x16 = (torch.randn(M, K, generator=gen) * 0.5).to(torch.float16)
        with torch.inference_mode():
            W = dequantize_gemm(qweight, qzeros, scales, 4, eff).clone()
            y16 = mod(x16).clone()
            ybf = mod(x16.to(torch.bfloat16)).clone()
        assert W.shape == (K, N) and W.dtype == torch.float16
        blobs[f"{name}.qweight"] = qweight.numpy()
        blobs[f"{name}.qzeros"] = qzeros.numpy()
        blobs[f"{name}.scales"] = scales.numpy()
        if bias:
            blobs[f"{name}.bias"] = b.numpy()
        blobs[f"{name}.x"] = x16.numpy()
        blobs[f"{name}.W"] = W.numpy()
        blobs[f"{name}.y_fp16"] = y16.numpy()
        blobs[f"{name}.y_bf16"] = ybf.float().numpy()
        meta[name] = dict(bits=4, group_size=gs, K=K, N=N, M=M, bias=bias)
        print(name, "W", tuple(W.shape), "y", tuple(y16.shape), float(y16.float().abs().mean()))
    blobs["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "awq_cases.npz"), **blobs)
    print("wrote awq_cases.npz")


if __name__ == "__main__":
    main()
