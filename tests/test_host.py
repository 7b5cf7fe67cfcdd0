"""CPU: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

import gptqmodel_b200 as g
from gptqmodel_b200 import B200QuantLinear

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b2q.h")).read()
    declared = set(re.findall(r"\b(b2q_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(g.SYMBOLS), (declared, set(g.SYMBOLS))
    raw = ctypes.CDLL(g.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert g.lib.b2q_version() == g.ABI_VERSION


def test_abi_argument_validation_without_gpu():
    assert g.lib.b2q_packed_bytes(4096, 4096, 4) == 4096 * 4096 // 2
    assert g.lib.b2q_packed_bytes(256, 128, 8) == 256 * 128
    assert g.lib.b2q_workspace_bytes(1, 4096, 4096, 1) == 0
    assert g.lib.b2q_workspace_bytes(16, 4096, 4096, 1) == 16 * 4096 * 2
    assert g.lib.b2q_workspace_bytes(16, 4096, 4096, 0) == 0
    assert g.lib.b2q_prepack(None, None, None, 64, 64, 4, None) == -2
    assert b"null pointer" in g.lib.b2q_last_error()
    one = ctypes.c_void_p(16)
    assert g.lib.b2q_prepack(one, None, one, 60, 64, 4, None) == -2
    assert b"K=60" in g.lib.b2q_last_error()
    assert g.lib.b2q_mm(one, one, one, None, None, None, one, 1, 64, 64, 3, 32, 0, None, 0, None) == -2
    assert b"bits=3" in g.lib.b2q_last_error()
    assert g.lib.b2q_mm(one, one, one, None, None, None, one, 1, 64, 64, 4, 48, 0, None, 0, None) == -2
    assert b"group_size=48" in g.lib.b2q_last_error()
    assert g.lib.b2q_mm(one, one, one, None, None, None, one, 0, 64, 64, 4, 32, 0, None, 0, None) == 0  # M == 0
    with pytest.raises(g.B2QError):
        g.check(-2, "x")


def test_validate_follows_reference_convention():
    ok, err = B200QuantLinear.validate(bits=3, group_size=128, in_features=128, out_features=128)
    assert not ok and isinstance(err, NotImplementedError)
    ok, err = B200QuantLinear.validate(bits=4, group_size=16, in_features=128, out_features=128)
    assert not ok and isinstance(err, NotImplementedError)
    ok, err = B200QuantLinear.validate(bits=4, group_size=128, in_features=100, out_features=128)
    assert not ok
    ok, err = B200QuantLinear.validate(bits=4, group_size=128, in_features=4096, out_features=4096,
                                       pack_dtype=torch.int32, dtype=torch.bfloat16)
    assert ok and err is None
    with pytest.raises(NotImplementedError):
        B200QuantLinear(bits=2, group_size=128, desc_act=False, sym=True, in_features=128, out_features=128)
    for attr in ("SUPPORTS_BACKENDS", "SUPPORTS_METHODS", "SUPPORTS_FORMATS", "SUPPORTS_BITS", "SUPPORTS_GROUP_SIZE",
                 "SUPPORTS_DESC_ACT", "SUPPORTS_SYM", "SUPPORTS_SHARDS", "SUPPORTS_TRAINING", "SUPPORTS_AUTO_PADDING",
                 "SUPPORTS_IN_FEATURES_DIVISIBLE_BY", "SUPPORTS_OUT_FEATURES_DIVISIBLE_BY", "SUPPORTS_PACK_DTYPES",
                 "SUPPORTS_ADAPTERS", "SUPPORTS_DEVICES", "SUPPORTS_PLATFORM", "SUPPORTS_DTYPES"):
        assert getattr(B200QuantLinear, attr) is not None, attr  # verify_supports_params (qlinear/__init__.py:300-332)
    assert B200QuantLinear.REQUIRES_FORMAT_V2 is True


def test_checkpoint_buffer_shapes_match_reference_layout():
    # tests/kernels/test_qlinear_hierarchy.py:239-267 — qweight [K*bits/32, N], qzeros [G, N*bits/32], scales [G, N]
    m = B200QuantLinear(bits=4, group_size=128, desc_act=False, sym=True, in_features=512, out_features=256, bias=True)
    assert m.qweight.shape == (512 * 4 // 32, 256) and m.qweight.dtype == torch.int32
    assert m.qzeros.shape == (4, 256 * 4 // 32) and m.scales.shape == (4, 256) and m.scales.dtype == torch.float16
    assert m.g_idx.shape == (512,) and m.g_idx[127] == 0 and m.g_idx[128] == 1
    assert m.bias.shape == (256,)
    m8 = B200QuantLinear(bits=8, group_size=-1, desc_act=False, sym=True, in_features=256, out_features=64)
    assert m8.qweight.shape == (64, 64) and m8.qzeros.shape == (1, 16) and m8.group_size == 256
    sd = m.state_dict()
    assert set(sd) == {"qweight", "qzeros", "scales", "g_idx", "bias"}  # the wire format


def test_no_cpu_fallback():
    m = B200QuantLinear(bits=4, group_size=128, desc_act=False, sym=True, in_features=128, out_features=64)
    with pytest.raises(g.B2QError):
        m.forward(torch.zeros(1, 128, dtype=torch.float16))  # before post_init
    with pytest.raises(g.B2QError):
        m.post_init()  # CPU tensors: must fail loudly, never fall back


def test_v1_to_v2_conversion():
    m = B200QuantLinear(bits=4, group_size=128, desc_act=False, sym=True, in_features=128, out_features=64)
    m.qzeros.data.fill_(0x77777777)
    m.qzero_format(1)
    m.convert_gptq_v1_to_v2()
    assert m.qzero_format() == 2 and int(m.qzeros[0, 0]) == 0x88888888 - (1 << 32)
