"""Property tests (hypothesis) of the integer re-packing code on the product side: random shapes / seeds, exact results.

  * pack_gptq == oracle.pack (which is pinned bit-exact to the reference's packers) for random grids and act-order maps,
  * AWQ -> GPTQ conversion: unpacking the converted tensors returns the codes that were packed the AWQ way,
  * v1 <-> v2 zero-point round trip through the module method,
  * tp shards: concatenating column shards / summing row-shard partial products reproduces the unsharded dequantised layer.
"""
import torch
from hypothesis import given, settings, strategies as st

import oracle
from gptqmodel_b200 import B200QuantLinear, awq_gemm_to_gptq, tp
from gptqmodel_b200.pack import pack_gptq

SET = settings(max_examples=25, deadline=None)


@SET
@given(bits=st.sampled_from([4, 8]), kg=st.integers(1, 4), gs=st.sampled_from([32, 64, 128]), n8=st.integers(1, 6),
       sym=st.booleans(), act=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_pack_gptq_equals_oracle(bits, kg, gs, n8, sym, act, seed):
    K, N = kg * gs, n8 * 8
    gen = torch.Generator().manual_seed(seed)
    W = torch.randn(N, K, generator=gen) * 0.4
    g_idx = torch.arange(K, dtype=torch.int32) // gs
    if act:
        g_idx = g_idx[torch.randperm(K, generator=gen)].contiguous()
    sc, ze, _ = (oracle.quantize_sym if sym else oracle.quantize_asym)(W, bits, gs, g_idx)
    a = pack_gptq(W, sc, ze, g_idx, bits)
    b = oracle.pack(W, sc, ze, g_idx, bits)
    assert torch.equal(a["qweight"], b[0]) and torch.equal(a["qzeros"], b[1]) and torch.equal(a["scales"], b[2])
    assert torch.equal(a["g_idx"], b[3])


@SET
@given(k8=st.integers(1, 32), g=st.sampled_from([8, 32, 64]), n8=st.integers(1, 8), seed=st.integers(0, 10 ** 6))
def test_awq_conversion_roundtrip(k8, g, n8, seed):
    K = k8 * 8
    if K % g:
        K = g * max(1, K // g)
    N = n8 * 8
    gen = torch.Generator().manual_seed(seed)
    codes = torch.randint(0, 16, (K, N), generator=gen)
    zeros = torch.randint(0, 16, (K // g, N), generator=gen)
    sc = (torch.rand(K // g, N, generator=gen) * 0.01 + 0.001).to(torch.float16)
    out = awq_gemm_to_gptq(oracle.awq_pack(codes), oracle.awq_pack(zeros), sc, g)
    assert torch.equal(oracle.unpack_qweight(out["qweight"], 4).to(torch.int64), codes)
    assert torch.equal(oracle.unpack_qzeros(out["qzeros"], 4).to(torch.int64), zeros)
    W = oracle.dequantize_weight(out["qweight"], out["qzeros"], out["scales"], out["g_idx"], 4)
    assert torch.equal(W, oracle.awq_dequantize(oracle.awq_pack(codes), oracle.awq_pack(zeros), sc, g))


@SET
@given(bits=st.sampled_from([4, 8]), seed=st.integers(0, 10 ** 6))
def test_v1_v2_zero_point_roundtrip(bits, seed):
    gen = torch.Generator().manual_seed(seed)
    m = B200QuantLinear(bits=bits, group_size=64, desc_act=False, sym=False, in_features=128, out_features=64)
    maxq = (1 << bits) - 1
    z = torch.randint(1, maxq + 1, (2, 64), generator=gen)       # true zero-points >= 1 (v1 stores z - 1 >= 0)
    pf = 32 // bits
    acc = torch.zeros(2, 64 // pf, dtype=torch.int64)
    for j in range(pf):
        acc |= z[:, j::pf] << (bits * j)
    v2 = torch.where(acc >= 2 ** 31, acc - 2 ** 32, acc).to(torch.int32)
    m.qzeros.data.copy_(oracle.convert_v2_to_v1(v2, bits))
    m.qzero_format(1)
    m.convert_gptq_v1_to_v2()
    assert m.qzero_format() == 2 and torch.equal(m.qzeros.data, v2)


@SET
@given(world=st.sampled_from([2, 4]), sym=st.booleans(), gs=st.sampled_from([32, 64]), seed=st.integers(0, 10 ** 4))
def test_tp_shards_recompose(world, sym, gs, seed):
    K, N = 256, 128
    from helpers import make_layer
    L = make_layer(K, N, group_size=gs, sym=sym, seed=seed)
    W = oracle.dequantize_weight(L["qweight"], L["qzeros"], L["scales"], L["g_idx"], 4).float()   # [K, N]
    x = torch.randn(3, K, generator=torch.Generator().manual_seed(seed))
    cols, part = [], torch.zeros(3, N)
    for r in range(world):
        c = tp.shard_columns(L, r, world)
        cols.append(oracle.dequantize_weight(c["qweight"], c["qzeros"], c["scales"], c["g_idx"], 4).float())
        rw = tp.shard_rows(L, r, world)
        Wr = oracle.dequantize_weight(rw["qweight"], rw["qzeros"], rw["scales"], rw["g_idx"], 4).float()
        part += x[:, r * K // world:(r + 1) * K // world] @ Wr
    assert torch.equal(torch.cat(cols, dim=1), W)
    assert torch.allclose(part, x @ W, atol=1e-3, rtol=1e-4)
