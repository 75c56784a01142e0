"""Row sharding of weight matrices over the GPUs of one box.

Quantize / dequantize / pack need no communication: ``W[out, in]`` is cut
along ``out`` and every rank quantizes its own rows — blockwise-64 blocks run
along ``in`` and never straddle a row, so the shards' codes, packed bytes and
scales are exactly the corresponding slices of the unsharded result.

The tensor-parallel linear is column-parallel: rank r holds the quantized rows
``[r0, r1)`` of W, computes ``y_r[M, r1-r0] = x @ dequant(W_r).T + b_r`` with the
same fused kernel, and the ranks' columns are assembled into ``y[M, out]`` —
the only exchange step on the path — in one of two ways:

* ``fused_gather=False``: ONE all-gather (NCCL over NVLink on the GPU box, gloo
  in the CPU tests) after the kernel;
* ``fused_gather=True`` (CUDA, one box): the GEMM's tile epilogue stores each
  finished tile straight into EVERY rank's ``y`` (symmetric memory): ONE TMA
  store to the buffer's NVSwitch multicast address, which the switch replicates
  to all ranks — or, without multicast (``fused_gather="peer"``), one store per
  peer-mapped buffer.  The gather overlaps the math tile by tile and the only
  thing after the kernel is a cross-rank barrier on the stream: one warp
  (``quanta_peer_barrier``) that releases its own dependents at once, so the
  next layer's weight stream starts while the ranks are still meeting.

One process per GPU; ``torch.distributed`` is plumbing only.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _host, _lib


def row_shard(n_rows, world_size, rank, multiple=1):
    """Rows ``[start, stop)`` of rank ``rank``: contiguous, sizes differ by at most
    ``multiple`` rows, every boundary a multiple of ``multiple`` (128 keeps GEMM
    tiles whole)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    units = -(-n_rows // multiple)
    base, extra = divmod(units, world_size)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    return min(u0 * multiple, n_rows), min(u1 * multiple, n_rows)


def shard_rows(weight, world_size, rank, multiple=1):
    """This rank's contiguous row slice of ``weight[out, in]`` (a view)."""
    r0, r1 = row_shard(weight.shape[0], world_size, rank, multiple)
    return weight[r0:r1]


def gather_columns(y_local, out_features, group=None, multiple=1):
    """All-gather the column shards ``y_local[M, r1-r0]`` of every rank into
    ``y[M, out_features]``.  Shards may differ in width (``row_shard``): they are
    padded to the widest one for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    if world == 1:
        return y_local
    M = y_local.shape[0]
    bounds = [row_shard(out_features, world, r, multiple) for r in range(world)]
    width = max(b - a for a, b in bounds)
    send = y_local
    if send.shape[1] != width:
        send = torch.zeros((M, width), dtype=y_local.dtype, device=y_local.device)
        send[:, : y_local.shape[1]] = y_local
    recv = torch.empty((world * M, width), dtype=y_local.dtype, device=y_local.device)      # rank-major concatenation
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    recv = recv.view(world, M, width)
    y = torch.empty((M, out_features), dtype=y_local.dtype, device=y_local.device)
    for r, (a, b) in enumerate(bounds):
        y[:, a:b] = recv[r, :, : b - a]
    return y


def x_device_type(device):
    """torch's DeviceType enum value for ``device`` (what ``has_multicast_support`` expects)."""
    from torch._C._autograd import DeviceType
    return DeviceType.CUDA if device.type == "cuda" else DeviceType.CPU


class TensorParallelLinear(nn.Module):
    """Column-parallel quantized linear over ``group``: this rank holds rows
    ``row_shard(out_features, world, rank, 128)`` of the weight, quantized
    blockwise along ``in_features`` (convention A), and ``forward`` all-gathers
    the ranks' output columns.  ``bits`` = 4 (packed) or 8."""

    def __init__(self, in_features, out_features, bits=4, bias=True, compute_dtype=torch.bfloat16, blocksize=64,
                 group=None, fused_gather=False, max_rows=256):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.bits, self.blocksize, self.compute_dtype, self.group = bits, blocksize, compute_dtype, group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rows = row_shard(out_features, self.world_size, self.rank, 128)
        self.register_buffer("qweight", None)
        self.register_buffer("scale", None)
        self.register_buffer("zero_point", None)
        self.register_buffer("bias", None)
        self._has_bias = bias
        # True: NVSwitch multicast from 4 ranks up when the fabric offers it (measured: 60 -> 43 us at M = 256 on 8
        # GPUs; with 2 ranks one store per peer is as fast or faster); "multicast" / "peer" force either
        self.fused_gather = fused_gather if (fused_gather and self.world_size > 1) else False
        self.max_rows = max_rows
        # True: for M <= 16 the ranks meet inside the GEMM kernel (quanta_gemm_wna16_scatter_sync) instead of in a barrier
        # kernel behind it.  Measured 1.5-2 us SLOWER (8 GPUs, M = 16: 17.9 vs 16.4 us; 2 GPUs: 21.3 vs 19.8): every CTA
        # pays a system-scope fence before the exit counter, where the kernel boundary flushes for free.  Off.
        self.kernel_sync = False
        self.own_barrier = True             # False: torch's symmetric-memory barrier kernel (A/B measurements)
        self._sym = None

    def _apply(self, fn, *args, **kwargs):
        """The kernels read scale / zero_point as float32: a ``.half()`` / ``.to(bfloat16)`` of the module must not
        cast them (same rule as quanta_b200.nn.linear)."""
        out = super()._apply(fn, *args, **kwargs)
        for name in ("scale", "zero_point"):
            buf = getattr(self, name, None)
            if isinstance(buf, torch.Tensor) and buf.dtype != torch.float32:
                setattr(self, name, buf.to(torch.float32))
        return out

    def _symmetric_outputs(self, device):
        """Two [max_rows, out_features] output buffers in symmetric memory, mapped into every
        rank of the group.  Two, because a peer may already be writing call n+1's tiles while
        this rank's consumers still read call n's y; the barrier of call n+1 orders call n+2's
        writes behind them (everything runs on the caller's stream)."""
        if self._sym is None:
            import torch.distributed._symmetric_memory as symm
            grp = self.group if self.group is not None else dist.group.WORLD
            buf = symm.empty((2, self.max_rows, self.out_features), dtype=self.compute_dtype, device=device)
            hdl = symm.rendezvous(buf, grp)
            mc = 0
            try:                                   # NVSwitch multicast mapping of the same buffer, when the fabric has one
                want = self.fused_gather == "multicast" or (self.fused_gather is True and self.world_size >= 4)
                if want and hdl.has_multicast_support(x_device_type(device), device.index):
                    mc = int(hdl.multicast_ptr)
            except Exception:                      # noqa: BLE001 - older torch: no multicast query
                mc = 0
            # completion flags for the in-kernel synchronisation of decode-sized batches
            # (quanta_gemm_wna16_scatter_sync): one unsigned int per peer, zero before the first call on every rank
            flags = symm.empty((max(self.world_size, 8),), dtype=torch.int32, device=device)
            flags.zero_()
            fhdl = symm.rendezvous(flags, grp)
            counter = torch.zeros(1, dtype=torch.int32, device=device)      # this rank's call counter (the kernel increments it)
            torch.cuda.synchronize(device)
            fhdl.barrier(channel=0)
            self._sym = {"buf": buf, "hdl": hdl, "ptrs": [int(p) for p in hdl.buffer_ptrs], "turn": 0, "mc": mc,
                         "flags": flags, "fhdl": fhdl, "flag_ptrs": [int(p) for p in fhdl.buffer_ptrs], "counter": counter, "epoch": 0}
        return self._sym

    @torch.no_grad()
    def load_shard(self, weight_rows, bias_rows=None):
        """Quantize this rank's rows (a CUDA tensor [r1-r0, in_features]); no communication."""
        from .functional.quantization import quantize_4bit, quantize_8bit
        r0, r1 = self.rows
        if tuple(weight_rows.shape) != (r1 - r0, self.in_features):
            raise ValueError(f"expected rows {r0}:{r1} of the weight, got shape {tuple(weight_rows.shape)}")
        if self.bits == 4:
            self.qweight, self.scale, self.zero_point = quantize_4bit(weight_rows, blocksize=self.blocksize, packed=True)
        else:
            self.qweight, self.scale, self.zero_point = quantize_8bit(weight_rows, blocksize=self.blocksize)
        if self._has_bias and bias_rows is not None:
            self.bias = bias_rows.detach().to(self.compute_dtype).clone()
        return self

    def local_matmul(self, x):
        """This rank's output columns ``y[:, r0:r1]`` — the fused dequant-GEMM kernel."""
        from .nn.functional import linear_wna16
        r0, r1 = self.rows
        return linear_wna16(x, self.qweight, self.scale, self.zero_point, self.bias, bits=self.bits,
                            blocksize=self.blocksize, out_features=r1 - r0, in_features=self.in_features)

    def forward(self, x):
        xin = x if x.dtype == self.compute_dtype else x.to(self.compute_dtype)
        lead = xin.shape[:-1]
        x2 = xin.reshape(-1, self.in_features)
        if self.fused_gather and x2.shape[0] <= self.max_rows:
            return self._forward_fused(x2).reshape(*lead, self.out_features)
        y_local = self.local_matmul(x2)
        y = gather_columns(y_local, self.out_features, self.group, 128)
        return y.reshape(*lead, self.out_features)

    def _forward_fused(self, x2):
        """GEMM with the gather in its epilogue.  Returns a view of the symmetric output buffer:
        valid until the call after next on this layer (copy it to keep it longer)."""
        from .nn.functional import linear_wna16_scatter
        sym = self._symmetric_outputs(x2.device)
        turn = sym["turn"]
        sym["turn"] = turn ^ 1
        M = x2.shape[0]
        slot = turn * self.max_rows * self.out_features * x2.element_size()
        r0, r1 = self.rows
        # one store per peer, or ONE store to the multicast address that the switch replicates to every rank
        targets = [sym["mc"] + slot] if sym["mc"] else [p + slot for p in sym["ptrs"]]
        # M <= 16: the ranks meet inside the kernel (its last CTA signals the peers and waits for them), so there is
        # no barrier kernel behind a GEMM that is itself only ~10 us long; otherwise one barrier on the stream
        sync = None
        if self.kernel_sync and M <= 16 and self.world_size <= 8:
            sym["epoch"] += 1                                       # host-side count of synchronised calls (tests)
            sync = (sym["flag_ptrs"], self.rank, self.world_size, sym["counter"].data_ptr())
        done = linear_wna16_scatter(x2, self.qweight, self.scale, self.zero_point, self.bias,
                                    (targets, self.out_features), r0, bits=self.bits,
                                    blocksize=self.blocksize, out_features=r1 - r0, sync=sync)
        if not done:
            # every rank's tiles have landed in this rank's buffer: one tiny kernel (quanta_peer_barrier) that lets the
            # NEXT layer's weight stream start while the ranks are still meeting
            if self.own_barrier and self.world_size <= 8:
                farr = (ctypes.c_void_p * len(sym["flag_ptrs"]))(*sym["flag_ptrs"])
                with _host.device_guard(x2.device):
                    st = _lib.lib().quanta_peer_barrier(farr, self.rank, self.world_size, sym["counter"].data_ptr(),
                                                        _host.stream_ptr(x2.device))
                _lib.check(st, "quanta_peer_barrier")
            else:
                sym["hdl"].barrier(channel=0)
        return sym["buf"][turn, :M]
