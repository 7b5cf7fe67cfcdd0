"""TEST INFRASTRUCTURE ONLY — CPU oracle for the GPTQ W4A16/W8A16 QuantLinear hot path.

A plain torch/numpy restatement (no code copied) of what the reference computes for
``gptqmodel.nn_modules.qlinear`` on a packed GPTQ layer.  Citations are
``/root/reference/<path>:<lines>``.

Parity pinning (see tests/test_oracle.py):
  * the reference's own 1024-value golden vector ``tests/q4_reference.py`` with the
    recipe of ``tests/test_q4_exllama_v2.py:32-87`` (committed as tests/golden/q4_reference.json),
  * the closed-form ``_reference_weight`` of ``tests/test_torch_kernel_accuracy.py:77-87``
    on that test's own generator (``_make_inputs`` :46-60),
  * v1<->v2 qzeros round trip of ``utils/model.py:810-818``.
The reference package itself cannot be imported in the authoring container (missing
pcre/logbar/device_smi/tokenicer/defuser/accelerate/torchao), so there are no
reference-generated fixtures beyond the golden vector above.

The product path (gptqmodel_b200/) never imports this module.
"""
from __future__ import annotations

import math

import numpy as np
import torch

__all__ = [
    "unpack_qweight",
    "unpack_qzeros",
    "dequantize_weight",
    "forward",
    "pack",
    "quantize_sym",
    "quantize_asym",
    "make_act_order",
    "convert_v1_to_v2",
    "convert_v2_to_v1",
    "CpuFusedLinear",
    "algorithmic_bytes",
]


# --------------------------------------------------------------------------------------
# unpack / dequant — nn_modules/qlinear/__init__.py:907-909 (shift tables), :962-975 (unpack),
# :1001-1003 (scale * (w - z)); eval fast path nn_modules/qlinear/torch.py:700-717.
# --------------------------------------------------------------------------------------
def _shifts(bits: int) -> torch.Tensor:
    return torch.arange(0, 32, bits, dtype=torch.int32)


def unpack_qweight(qweight: torch.Tensor, bits: int) -> torch.Tensor:
    """int32 [K*bits/32, N] -> integer codes [K, N] (int16).

    Word [i, n] holds rows i*pf .. i*pf+pf-1 of column n, row i*pf+j in bits [bits*j, bits*(j+1)).
    (qlinear/__init__.py:968-975: expand over pack_factor, right-shift by wf.unsqueeze(-1), & maxq.)
    Right shift of negative int32 is arithmetic in torch; the mask removes the sign fill.
    """
    assert qweight.dtype == torch.int32
    pf = 32 // bits
    maxq = (1 << bits) - 1
    sh = _shifts(bits).view(1, pf, 1)
    w = torch.bitwise_right_shift(qweight.unsqueeze(1).expand(-1, pf, -1), sh)
    w = torch.bitwise_and(w, maxq).to(torch.int16)
    return w.reshape(qweight.shape[0] * pf, qweight.shape[1])


def unpack_qzeros(qzeros: torch.Tensor, bits: int) -> torch.Tensor:
    """int32 [G, N*bits/32] -> zero codes [G, N] (int16), v2 semantics (true zero-point).

    Word [g, c] holds columns c*pf .. c*pf+pf-1, column c*pf+j in bits [bits*j, ..)
    (qlinear/__init__.py:962-966; torch.py:465-478).
    """
    assert qzeros.dtype == torch.int32
    pf = 32 // bits
    maxq = (1 << bits) - 1
    sh = _shifts(bits).view(1, 1, pf)
    z = torch.bitwise_right_shift(qzeros.unsqueeze(2).expand(-1, -1, pf), sh)
    z = torch.bitwise_and(z, maxq).to(torch.int16)
    return z.reshape(qzeros.shape[0], qzeros.shape[1] * pf)


def dequantize_weight(qweight, qzeros, scales, g_idx, bits: int) -> torch.Tensor:
    """W[k, n] = scales[g_idx[k], n] * (q[k, n] - z[g_idx[k], n]) in ``scales.dtype``.

    qlinear/__init__.py:1001-1003: integer subtract first, then ONE rounding in the multiply
    (int16 operand promotes to the float dtype exactly; |q-z| <= 255).
    """
    w = unpack_qweight(qweight, bits)
    z = unpack_qzeros(qzeros, bits)
    gi = g_idx.long()
    return scales[gi] * (w - z[gi])


def forward(x, qweight, qzeros, scales, g_idx, bits: int, bias=None, accumulate_fp32: bool = True):
    """out = x.reshape(-1, K) @ W.to(x.dtype) (+ bias) — qlinear/torch.py:302-347.

    The reference calls torch.matmul in x.dtype (cuBLAS: fp32 accumulate, one final rounding).
    On CPU we compute the product in fp32 from the SAME rounded operands and round once at the
    end, which is the arithmetic cuBLAS performs (accumulate_fp32=True, default).
    """
    out_shape = x.shape[:-1] + (qweight.shape[1],)
    x2 = x.reshape(-1, x.shape[-1])
    W = dequantize_weight(qweight, qzeros, scales, g_idx, bits).to(x.dtype)
    if accumulate_fp32:
        out = (x2.float() @ W.float())
        if bias is not None:
            # reference adds bias in the output dtype AFTER rounding the matmul (torch.py:338-342)
            out = out.to(x.dtype)
            out = out + bias.to(x.dtype)
        out = out.to(x.dtype)
    else:
        out = x2 @ W
        if bias is not None:
            out = out + bias.to(x.dtype)
    return out.reshape(out_shape)


# --------------------------------------------------------------------------------------
# pack — restates pack_original, qlinear/__init__.py:1500-1583 (bits 2/4/8 branch).
# --------------------------------------------------------------------------------------
def pack(weight: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, g_idx: torch.Tensor, bits: int):
    """weight [N, K] float, scales/zeros [N, G] -> (qweight int32 [K*bits/32, N],
    qzeros int32 [G, N*bits/32] (v2: true zero), scales fp16 [G, N], g_idx int32 [K]).

    q = clamp(round((W + z*s)/s), 0, maxq)   (:1529-1531), using the fp32 scales,
    then rows packed LSB-first into int32 words (:1536-1539), zeros packed along N (:1580-1583).
    """
    assert bits in (2, 4, 8)
    pf = 32 // bits
    maxq = (1 << bits) - 1
    W = weight.float()
    s = scales.float().T.contiguous()  # [G, N]
    z = zeros.float().T.contiguous()  # [G, N]
    gi = g_idx.long()
    sz = z * s
    q = torch.round((W + sz[gi].T) / s[gi].T).clamp_(0, maxq)  # [N, K]
    q = q.to(torch.int64).T.contiguous().numpy().astype(np.uint32)  # [K, N]
    K, N = q.shape
    assert K % pf == 0 and N % pf == 0
    qweight = np.zeros((K // pf, N), dtype=np.uint32)
    for j in range(pf):
        qweight |= q[j::pf] << np.uint32(bits * j)
    zi = z.numpy().astype(np.uint32)
    G = zi.shape[0]
    qzeros = np.zeros((G, N // pf), dtype=np.uint32)
    for j in range(pf):
        qzeros |= zi[:, j::pf] << np.uint32(bits * j)
    return (
        torch.from_numpy(qweight.view(np.int32)),
        torch.from_numpy(qzeros.view(np.int32)),
        s.to(torch.float16),
        g_idx.to(torch.int32).clone(),
    )


# --------------------------------------------------------------------------------------
# quantiser fixtures — tests/kernels/test_swordfish.py:31-56 (sym) and a min/max asym twin
# (zero = round(-min/scale) clamp 0..maxq).
# --------------------------------------------------------------------------------------
def quantize_sym(weight: torch.Tensor, bits: int, group_size: int, g_idx: torch.Tensor | None = None):
    """Returns (scales [N, G], zeros [N, G], g_idx [K]) — sym grid, zero = 2^(bits-1)."""
    N, K = weight.shape
    half = 1 << (bits - 1)
    if g_idx is None:
        g_idx = torch.arange(K, dtype=torch.int32) // group_size
    G = int(g_idx.max().item()) + 1
    scales = torch.zeros((N, G), dtype=weight.dtype)
    for g in range(G):
        blk = weight[:, g_idx == g]
        m = blk.abs().max(dim=1, keepdim=True).values
        m[m == 0] = 1.0
        scales[:, g : g + 1] = m / (half - 1)
    zeros = torch.full((N, G), float(half), dtype=weight.dtype)
    return scales, zeros, g_idx


def quantize_asym(weight: torch.Tensor, bits: int, group_size: int, g_idx: torch.Tensor | None = None):
    N, K = weight.shape
    maxq = (1 << bits) - 1
    if g_idx is None:
        g_idx = torch.arange(K, dtype=torch.int32) // group_size
    G = int(g_idx.max().item()) + 1
    scales = torch.zeros((N, G), dtype=weight.dtype)
    zeros = torch.zeros((N, G), dtype=weight.dtype)
    for g in range(G):
        blk = weight[:, g_idx == g].float()
        lo = blk.min(dim=1, keepdim=True).values.clamp(max=0)
        hi = blk.max(dim=1, keepdim=True).values.clamp(min=0)
        sc = (hi - lo) / maxq
        sc[sc == 0] = 1.0
        scales[:, g : g + 1] = sc.to(weight.dtype)
        zeros[:, g : g + 1] = torch.round(-lo / sc).clamp(0, maxq).to(weight.dtype)
    return scales, zeros, g_idx


def make_act_order(K: int, group_size: int, seed: int = 42):
    """perm + g_idx as in tests/kernels/test_swordfish.py:127-131: g_idx = (arange//g)[perm]."""
    gen = torch.Generator().manual_seed(seed)
    perm = torch.randperm(K, generator=gen)
    g_idx = (torch.arange(K, dtype=torch.int32) // group_size)[perm]
    return perm, g_idx


# --------------------------------------------------------------------------------------
# qzeros v1 <-> v2 — utils/model.py:810-818 (+0x11111111 for 4-bit, +0x01010101 for 8-bit).
# --------------------------------------------------------------------------------------
_V1_OFFSET = {2: 0x55555555, 4: 0x11111111, 8: 0x01010101}


def convert_v1_to_v2(qzeros: torch.Tensor, bits: int) -> torch.Tensor:
    a = qzeros.numpy().view(np.uint32) + np.uint32(_V1_OFFSET[bits])
    return torch.from_numpy(a.view(np.int32).copy())


def convert_v2_to_v1(qzeros: torch.Tensor, bits: int) -> torch.Tensor:
    a = qzeros.numpy().view(np.uint32) - np.uint32(_V1_OFFSET[bits])
    return torch.from_numpy(a.view(np.int32).copy())


# --------------------------------------------------------------------------------------
# CPU fused baseline — TorchAtenLinear: qlinear/torch_aten_kernel.py:32-42 (zero offsets),
# :128-158 (_build_ret_idx), :184-209 (transform_cpu), :247-263 (_fused_op_forward);
# pack_scales_and_zeros qlinear/torch_fused.py:24-37.  The arithmetic itself lives in PyTorch
# ATen (torch>=2.8 per the reference's requirements.txt; 2.11.0 here):
# aten::_convert_weight_to_int4pack_for_cpu / aten::_weight_int4pack_mm_for_cpu.
# No synthetic KAT exists in the reference for this op ("parity unpinned" at that boundary);
# tests/test_oracle.py pins it against `forward` above at bf16 tolerance.
# --------------------------------------------------------------------------------------
class CpuFusedLinear:
    def __init__(self, qweight, qzeros, scales, g_idx, bits: int, group_size: int, bias=None):
        assert bits == 4, "aten int4pack path is 4-bit only"
        K = g_idx.shape[0]
        self.group_size = group_size if group_size > 0 else K
        self.bias = bias
        sc = scales.to(torch.bfloat16).contiguous()
        w = unpack_qweight(qweight, bits).to(torch.uint8)  # [K, N]
        self.ret_idx = self._build_ret_idx(g_idx, self.group_size)
        w = w.index_select(0, self.ret_idx.long()).t().contiguous()  # [N, K]
        self.qweight = torch.ops.aten._convert_weight_to_int4pack_for_cpu(w.int(), 1).contiguous()
        zc = unpack_qzeros(qzeros, bits)
        zoff = ((1 << (bits - 1)) - zc.to(sc.dtype)) * sc  # torch_aten_kernel.py:32-42
        self.scales_and_zeros = torch.cat([sc.unsqueeze(2), zoff.unsqueeze(2)], 2).contiguous()
        self.identity_perm = bool(torch.equal(self.ret_idx, torch.arange(K, dtype=torch.int32)))

    @staticmethod
    def _build_ret_idx(g_idx: torch.Tensor, group_size: int) -> torch.Tensor:
        # rows sorted by group, stable inside a group (torch_aten_kernel.py:128-158)
        total = g_idx.shape[0]
        g = g_idx.to(torch.int64)
        pos = torch.zeros(total, dtype=torch.int64)
        order = torch.argsort(g, stable=True)
        sorted_g = g[order]
        start = torch.searchsorted(sorted_g, sorted_g, right=False)
        rank_in_group = torch.arange(total) - start
        pos[order] = sorted_g * group_size + rank_in_group
        ret = torch.zeros(total, dtype=torch.int32)
        ret[pos] = torch.arange(total, dtype=torch.int32)
        return ret

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out_shape = x.shape[:-1] + (self.scales_and_zeros.shape[1],)
        x2 = x.reshape(-1, x.shape[-1])
        if not self.identity_perm:
            x2 = x2[:, self.ret_idx.long()]
        x2 = x2.contiguous()
        od = x2.dtype
        if od != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)
        out = torch.ops.aten._weight_int4pack_mm_for_cpu(x2, self.qweight, self.group_size, self.scales_and_zeros)
        if od != torch.bfloat16:
            out = out.to(od)
        out = out.reshape(out_shape)
        if self.bias is not None:
            out = out + self.bias.to(out.dtype)
        return out


# --------------------------------------------------------------------------------------
# algorithmic bytes per QuantLinear call — SURVEY.md §8(d) / BASELINE.md §2.
# --------------------------------------------------------------------------------------
def algorithmic_bytes(K: int, N: int, group_size: int, bits: int, M: int, desc_act: bool = False, bias: bool = False) -> int:
    g = group_size if group_size > 0 else K
    G = math.ceil(K / g)
    b = K * N * bits // 8 + G * N * 2 + G * (N * bits // 32) * 4 + M * K * 2 + M * N * 2
    if desc_act:
        b += 4 * K
    if bias:
        b += 2 * N
    return b
