"""Mixture-of-experts block over QuantLinear experts (BASELINE configs[4]: Mixtral-8x7B int4 g64 asym).

In the reference every expert's ``w1 / w3 / w2`` is an independent QuantLinear module
(/root/reference/gptqmodel/models/definitions/mixtral.py:30-34) that the model's own Python loop calls per expert; the
fused MoE kernel it ships (``swordfish_moe.cu``) is exported but never called.  This module is that
per-expert loop, arranged for the B200 kernels and for tensor parallelism:

  * tokens are sorted by expert ONCE, every expert then sees one contiguous block of its routed tokens — a plain batched
    GEMM of the right tier for its block size (decode tier for <= 8 tokens, tcgen05 tiers above);
  * ``w1`` and ``w3`` of an expert consume the same block: with `fuse_siblings` they are ONE decode launch;
  * tensor parallel: ``w1 / w3`` column-sharded, ``w2`` row-sharded (`tp.shard_moe_expert`), so every rank holds a slice
    of EVERY expert and the block ends in exactly one all-reduce of the combined output, as for a dense MLP.

Experts are any callables mapping ``[m, K] -> [m, N]`` (B200QuantLinear modules in production; dense stand-ins in the
CPU tests), so the routing / combination / sharding algebra is testable without a GPU.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.nn.functional as F

from . import tp


def route_topk(router_logits: torch.Tensor, top_k: int):
    """Mixtral routing: softmax over experts, top-k, renormalise (HF MixtralSparseMoeBlock)."""
    probs = F.softmax(router_logits.float(), dim=-1)
    w, ids = torch.topk(probs, top_k, dim=-1)
    w = w / w.sum(dim=-1, keepdim=True)
    return ids, w


class MoEExperts(torch.nn.Module):
    """``y = sum_k w_k * w2_e( silu(w1_e x) * w3_e x )`` over the top-k experts e of every token."""

    def __init__(self, w1: Sequence[Callable], w3: Sequence[Callable], w2: Sequence[Callable], fuse: bool = True,
                 group=None, reduce=None):
        super().__init__()
        if not (len(w1) == len(w3) == len(w2)) or len(w1) == 0:
            raise ValueError("MoEExperts: need the same number (>= 1) of w1 / w3 / w2 experts")
        as_list = lambda xs: torch.nn.ModuleList(xs) if all(isinstance(x, torch.nn.Module) for x in xs) else list(xs)  # noqa: E731
        self.w1, self.w3, self.w2 = as_list(w1), as_list(w3), as_list(w2)
        self.group = group
        self.reduce = reduce  # optional tp.P2PAllReduce for decode-sized outputs
        if fuse:
            from .qlinear import B200QuantLinear, fuse_siblings

            for a, b in zip(self.w1, self.w3):
                if isinstance(a, B200QuantLinear) and isinstance(b, B200QuantLinear):
                    fuse_siblings([a, b])

    @property
    def num_experts(self) -> int:
        return len(self.w1)

    def forward(self, x: torch.Tensor, topk_ids: torch.Tensor, topk_weights: torch.Tensor) -> torch.Tensor:
        """x [T, K]; topk_ids / topk_weights [T, top_k] (weights already normalised) -> [T, K_out] (all-reduced)."""
        T, top_k = topk_ids.shape
        flat_e = topk_ids.reshape(-1)
        order = torch.argsort(flat_e, stable=True)             # (token, k) pairs sorted by expert
        tok = order // top_k
        counts = torch.bincount(flat_e, minlength=self.num_experts).tolist()  # one host sync per block, like the
        xs = x.index_select(0, tok)                                            # reference's per-expert Python loop
        wts = topk_weights.reshape(-1).index_select(0, order).to(torch.float32)
        out = None
        start = 0
        for e, cnt in enumerate(counts):
            if cnt == 0:
                continue
            blk = xs[start:start + cnt]
            h = F.silu(self.w1[e](blk)) * self.w3[e](blk)
            y = self.w2[e](h)
            if out is None:
                out = torch.zeros((T, y.shape[-1]), dtype=torch.float32, device=x.device)
            out.index_add_(0, tok[start:start + cnt], y.float() * wts[start:start + cnt, None])
            start += cnt
        if out is None:
            raise ValueError("MoEExperts.forward: empty routing")
        out = out.to(x.dtype)
        if self.reduce is not None and out.numel() <= self.reduce.max_elems and out.numel() % 8 == 0:
            return self.reduce(out.contiguous())
        return tp.all_reduce_sum_(out, self.group)
