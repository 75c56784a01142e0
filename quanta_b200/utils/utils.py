"""Drop-in CUDA implementation of the 4-bit pack helpers of ``Quanta.utils``
(Quanta/utils/utils.py:23-54; code-identical copy in utils/tensor_utils.py:6-37)."""
from __future__ import annotations

import torch

from .. import _host, _lib


def pack_4bit_tensor(tensor):
    """Pack 4-bit values into a tensor with half the size.  Returns
    ``(packed flat uint8 of ceil(numel/2), origin_shape)`` like utils.py:23-35:
    even index -> low nibble, one zero pad if numel is odd, inputs > 15 are not
    masked."""
    if tensor.dtype != torch.uint8:
        raise ValueError("Input tensor must be uint8")
    _host.require_cuda(tensor)
    origin_shape = tensor.shape
    q = tensor.detach().reshape(-1)
    if not q.is_contiguous():
        q = q.contiguous()
    n = q.numel()
    packed = torch.empty((n + 1) // 2, dtype=torch.uint8, device=q.device)
    if n:
        with torch.cuda.device(q.device):
            st = _lib.lib().quanta_pack4(q.data_ptr(), n, packed.data_ptr(), _host.stream_ptr(q.device))
        _lib.check(st, "quanta_pack4")
    return packed, origin_shape


def unpack_4bit_tensor(packed_tensor):
    """Unpack a tensor where each byte contains two 4-bit values; returns the
    flat uint8 tensor of 2*len (pad element included; utils.py:37-48)."""
    _host.require_cuda(packed_tensor, "packed_tensor")
    p = packed_tensor.detach().reshape(-1)
    if p.dtype != torch.uint8:
        p = p.to(torch.uint8)
    if not p.is_contiguous():
        p = p.contiguous()
    out = torch.empty(p.numel() * 2, dtype=torch.uint8, device=p.device)
    if p.numel():
        with torch.cuda.device(p.device):
            st = _lib.lib().quanta_unpack4(p.data_ptr(), p.numel(), out.data_ptr(), _host.stream_ptr(p.device))
        _lib.check(st, "quanta_unpack4")
    return out


def tensor_bits_to_bytes(tensor, bits):
    """Convert tensor size from bits to bytes (utils.py:50-54)."""
    n = 1
    for d in tensor.shape:
        n *= int(d)
    return (n * bits + 7) // 8
