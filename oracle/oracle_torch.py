"""CPU restatement of the reference's hot path as the chain of eager torch ops the
reference itself runs (TEST INFRASTRUCTURE / CPU BASELINE ONLY — nothing under
quanta_b200/ may import this module).

oracle_np.py / quanta_oracle.c restate the *arithmetic* one rounding at a time;
this module restates the *implementation*: the same sequence of whole-tensor
torch operations, each a separate pass with its own temporary, threaded by ATen's
intra-op pool.  That is what a user of the reference measures on the host cores,
so `bench.py --impl reference` times this one (the reference package itself does
not exist on the GPU box).  Results are checked bit-for-bit against oracle_np in
tests/test_oracle_torch.py.

  quantize_linear      Quanta/functional/quantization.py:73-99 (4-bit) / :185-210 (8-bit):
                       min/max (all elements or dim 0, keepdim), degenerate-range fix,
                       scale = (max - min) / L, zero_point = min,
                       q = clamp(round((x - min) / scale), 0, L).to(uint8)
  quantize_block       the per_channel branch applied to x.reshape(-1, B).t()  
  pack4                Quanta/utils/utils.py:23-35
  dequantize_linear    Quanta/functional/quantization.py:38 / :58
"""
from __future__ import annotations

import torch


def quantize_linear(x, bits=8, per_channel=False):
    levels = 255 if bits == 8 else 15
    if per_channel:
        mn = x.min(dim=0, keepdim=True)[0]
        mx = x.max(dim=0, keepdim=True)[0]
        mask = mx == mn
        mx = torch.where(mask, mn + 1e-6, mx)
    else:
        mn, mx = x.min(), x.max()
        if mx == mn:
            mx = mn + 1e-6
    scale = (mx - mn) / levels
    q = torch.clamp(torch.round((x - mn) / scale), 0, levels).to(torch.uint8)
    return q, scale, mn


def quantize_block(x, bits=4, block=64):
    """Blockwise = per_channel on the [block, n/block] transposed view; codes come back in flat order."""
    q, scale, zp = quantize_linear(x.reshape(-1, block).t(), bits, per_channel=True)
    return q.t().reshape(x.shape), scale.reshape(-1), zp.reshape(-1)


def pack4(q):
    flat = q.flatten()
    if flat.numel() % 2:
        flat = torch.cat([flat, torch.zeros(1, dtype=torch.uint8)])
    pairs = flat.reshape(-1, 2)
    return pairs[:, 0] | (pairs[:, 1] << 4)


def quantize4_block_pack(x, block=64):
    q, scale, zp = quantize_block(x, 4, block)
    return pack4(q), scale, zp


def dequantize_linear(q, scale, zp):
    return q.float() * scale + zp
