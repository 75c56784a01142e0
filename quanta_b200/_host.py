"""Shared host-side plumbing: device checks, dtype codes, stream, workspace."""
from __future__ import annotations

import torch

from . import _lib

_DTYPE = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}
_workspaces = {}


def require_cuda(t, name="tensor"):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"quanta_b200 is the CUDA backend of this path and has no CPU fallback: "
                           f"{name} is on {t.device}")


def dtype_code(t):
    try:
        return _DTYPE[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; expected float32, float16 or bfloat16") from None


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device):
    """cudaStream_t of torch's current stream on ``device`` (the raw getter costs ~0.3 us; the Stream object ~2 us)."""
    if _raw_stream is not None:
        idx = device.index
        return _raw_stream(torch.cuda.current_device() if idx is None else idx)
    return torch.cuda.current_stream(device).cuda_stream


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """``torch.cuda.device(device)`` only when it would change anything (the context manager costs ~4 us per call)."""
    if device.index is None or torch.cuda.current_device() == device.index:
        return _NO_GUARD
    return torch.cuda.device(device)


_ws_bytes_cache = {}


def workspace_bytes(op, rows, cols):
    """``quanta_workspace_bytes`` memoised per (op, rows, cols): a pure function, asked for on every call."""
    key = (op, rows, cols)
    v = _ws_bytes_cache.get(key)
    if v is None:
        v = int(_lib.lib().quanta_workspace_bytes(op, rows, cols))
        if len(_ws_bytes_cache) > 4096:
            _ws_bytes_cache.clear()
        _ws_bytes_cache[key] = v
    return v


def workspace(device, nbytes):
    """Per (device, stream) scratch buffer, grown on demand.  Reuse is safe
    because every kernel that touches it is ordered on that stream."""
    key = (device.index, stream_ptr(device))
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


_quantize_workspaces = {}
_ERRPTR_BYTES = slice(48, 56)          # include/quanta_b200.h: address of the host-mapped error flag


def quantize_workspace(device, nbytes):
    """Per (device, stream) workspace of the quantize entries.  Its header holds the grid-barrier
    counters of the single-launch per-tensor / per-channel kernels, which must be zero before the
    first call and are left zero by every call (include/quanta_b200.h): zero-initialised once, never
    shared with other kernels' scratch.  A pinned host int32 is registered in the header as the
    kernels' error flag: if a grid barrier timed out in an earlier call on this workspace, the header
    is re-zeroed and QuantaError is raised here, before the next launch."""
    key = (device.index, stream_ptr(device))
    rec = _quantize_workspaces.get(key)
    if rec is not None and int(rec["flag"][0]) != 0:
        rec["flag"].zero_()
        rec["buf"][:256].zero_()
        rec["buf"][_ERRPTR_BYTES].view(torch.int64).fill_(rec["flag"].data_ptr())
        raise _lib.QuantaError("an earlier quantize call on this stream timed out at its grid barrier (its outputs are "
                               "undefined); the workspace header has been reset — repeat the call")
    if rec is None or rec["buf"].numel() < nbytes:
        flag = rec["flag"] if rec is not None else torch.zeros(1, dtype=torch.int32).pin_memory()
        buf = torch.zeros(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        buf[_ERRPTR_BYTES].view(torch.int64).fill_(flag.data_ptr())      # pinned host memory is device-visible (UVA)
        rec = {"buf": buf, "flag": flag}
        _quantize_workspaces[key] = rec
    return rec["buf"]


_gemm_workspaces = {}


def gemm_workspace(device, nbytes):
    """Per (device, stream) workspace of the dequant-GEMM.  Its head holds the
    stream-K arrival counters, which must be zero before the first call and are
    left zero by every call (include/quanta_b200.h), so this buffer is
    zero-initialised once and never shared with the other kernels' scratch."""
    key = (device.index, stream_ptr(device))
    buf = _gemm_workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
        _gemm_workspaces[key] = buf
    return buf


def rows_cols(t):
    """[rows, cols] view of a contiguous tensor: dim 0 x everything else."""
    if t.dim() == 0:
        return 1, 1
    rows = t.shape[0]
    return rows, (t.numel() // rows if rows else 0)
