"""Model-level drop-in (the caller side of the hot path): every nn.Linear of a tiny HF Llama is replaced by a quantised
module and the logits are compared with the same model holding the DEQUANTISED dense weights.

On the CPU the oracle stands in for the kernel (this checks the swap walk, the RTN grid, the packer and the harness
itself); the GPU twin of this test (tests/test_gpu_parity_formats.py) runs the real B200QuantLinear modules.
"""
import pytest
import torch
import torch.nn as nn

import oracle
from gptqmodel_b200 import convert


class OracleLinear(nn.Module):
    """CPU stand-in with the QuantLinear forward contract, computing through the oracle."""

    def __init__(self, q):
        super().__init__()
        self.q = q

    def forward(self, x):
        q = self.q
        y = oracle.forward(x.reshape(-1, x.shape[-1]), q.qweight.data, q.qzeros.data, q.scales.data, q.g_idx.data, q.bits,
                           bias=None if q.bias is None else q.bias.data)
        return y.reshape(x.shape[:-1] + (y.shape[-1],))


def tiny_llama(dtype=torch.float16):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=1000, max_position_embeddings=128)
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg).eval().to(dtype)


def test_rtn_grid_bounds_the_quantisation_error():
    gen = torch.Generator().manual_seed(0)
    W = torch.randn(64, 256, generator=gen) * 0.3
    for bits, gs, sym in ((4, 128, True), (4, 64, False), (8, 32, False), (4, -1, True)):
        lin = nn.Linear(256, 64, bias=False)
        lin.weight.data.copy_(W)
        q = convert.quantize_linear(lin, bits=bits, group_size=gs, sym=sym, device="cpu")
        Wq = oracle.dequantize_weight(q.qweight.data, q.qzeros.data, q.scales.data, q.g_idx.data, bits).float().T  # [N, K]
        step = q.scales.data.float().T.repeat_interleave(gs if gs > 0 else 256, dim=1)                             # [N, K]
        assert ((Wq - W).abs() <= 0.51 * step + 1e-3).all()        # round-to-nearest: within half a step (+ fp16 scales)
    with pytest.raises(ValueError):
        convert.rtn_grid(torch.zeros(8, 100), 4, 64, True)


def test_replace_linears_in_a_tiny_llama_matches_dense_dequantised_model():
    model = tiny_llama(torch.float32)   # fp32 activations keep the CPU run fast; the stand-in rounds like the kernel
    dense = tiny_llama(torch.float32)
    quantised = {}

    def factory(name, lin):
        lin16 = nn.Linear(lin.in_features, lin.out_features, bias=lin.bias is not None)
        lin16.weight.data.copy_(lin.weight.data)
        q = convert.quantize_linear(lin16, bits=4, group_size=128, sym=("mlp" in name), device="cpu", name=name)
        quantised[name] = q
        return OracleLinear(q)

    swapped = convert.replace_linears(model, factory)
    assert len(swapped) == 14 and "lm_head" not in swapped and all(isinstance(m, OracleLinear) for m in swapped.values())
    assert isinstance(model.model.layers[1].mlp.down_proj, OracleLinear) and isinstance(model.lm_head, nn.Linear)
    for name, q in quantised.items():   # the dense twin gets exactly the dequantised weights
        W = oracle.dequantize_weight(q.qweight.data, q.qzeros.data, q.scales.data, q.g_idx.data, 4).float().T
        dense.get_submodule(name).weight.data.copy_(W)
    ids = torch.randint(0, 1000, (2, 9), generator=torch.Generator().manual_seed(3))
    with torch.inference_mode():
        a = model(ids).logits
        b = dense(ids).logits
    assert a.shape == (2, 9, 1000)
    assert (a - b).abs().max().item() < 2e-2 * b.abs().max().item()
    kept = convert.replace_linears(tiny_llama(torch.float32), lambda n, l: None)
    assert kept == {}
