"""Tensor-parallel dequant-GEMM rows of bench.py's N > 1 line (BASELINE.json configs[4]):
column-parallel W4A16 linear on Llama-3-70B shapes — local kernel, NCCL all-gather, and the gather fused into
the GEMM epilogue (symmetric memory; NVSwitch multicast from 4 ranks up) — timed as device time under CUDA-graph
replay (max over ranks) and checked in the same run:

  * gather parity (bit-exact): every rank's y[:, r0:r1] equals its own local kernel output, the NCCL and the
    fused paths agree, and all ranks hold the same y (checksums compared across ranks);
  * layer parity: y against the single-device layer computed from the SAME full weight on this rank
    (`bit_identical` is informational: a different row count changes the stream-K split and with it the
    fp32 summation order; `rel_err` is held to the 1e-2 tolerance of rows G1/G2)."""
import torch


def tp_rows(device, rank, world, pk, torch_mod, dist, Q, reps=20):
    from quanta_b200.nn import linear_wna16
    from quanta_b200.sharding import TensorParallelLinear

    def max_over_ranks(v):
        t = torch.tensor([v], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(flag):
        t = torch.tensor([1 if flag else 0], device=device, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def graph_us(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(); dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                y = fn()
        g.replay(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = max_over_ranks(e0.elapsed_time(e1) * 1e3 / reps)
        del g, y
        return us

    rows = []
    for (No, Ki) in ((8192, 8192), (28672, 8192)):
        gen = torch.Generator(device=device).manual_seed(777 + No)          # the same full weight on every rank
        w = torch.randn(No, Ki, device=device, generator=gen) * 0.02
        qf = Q.quantize_4bit(w, blocksize=64, packed=True)
        lin = TensorParallelLinear(Ki, No, bits=4, bias=False, compute_dtype=torch.bfloat16)
        r0, r1 = lin.rows
        lin.load_shard(w[r0:r1].contiguous())
        linf = TensorParallelLinear(Ki, No, bits=4, bias=False, compute_dtype=torch.bfloat16, fused_gather=True)
        linf.qweight, linf.scale, linf.zero_point = lin.qweight, lin.scale, lin.zero_point
        del w
        for M in (16, 256):
            xg = torch.Generator(device=device).manual_seed(99 + M)
            x = torch.randn(M, Ki, device=device, generator=xg).to(torch.bfloat16)
            # ---- parity
            y_local = lin.local_matmul(x)
            y_nccl = lin(x)
            y_fused = linf(x).clone()
            ref = linear_wna16(x, *qf, None, bits=4, blocksize=64, out_features=No)
            torch.cuda.synchronize()
            gather_ok = bool(torch.equal(y_nccl[:, r0:r1], y_local)) and bool(torch.equal(y_fused, y_nccl))
            chk = y_fused.view(torch.int16).to(torch.int64).sum().reshape(1)
            chks = [torch.zeros_like(chk) for _ in range(world)]
            dist.all_gather(chks, chk)
            same_everywhere = all(int(c.item()) == int(chks[0].item()) for c in chks)
            rel = float((y_fused.float() - ref.float()).abs().max() / ref.float().abs().max())
            bit_layer = bool(torch.equal(y_fused, ref))
            # ---- device time
            us_local = graph_us(lambda: lin.local_matmul(x))
            us_nccl = graph_us(lambda: lin(x))
            us_fused = graph_us(lambda: linf(x))
            # the same with the ranks' completion inside the kernel (M <= 16; off by default: it measures slower)
            us_fused_kernel_sync = None
            if M <= 16:
                linf.kernel_sync = True
                us_fused_kernel_sync = graph_us(lambda: linf(x))
                linf.kernel_sync = False
            # the same with torch's symmetric-memory barrier kernel instead of quanta_peer_barrier
            linf.own_barrier = False
            us_fused_torch_barrier = graph_us(lambda: linf(x))
            linf.own_barrier = True
            flops = 2.0 * M * No * Ki
            sent = M * (r1 - r0) * 2
            rows.append({"op": "tp_linear W4A16", "N": No, "K": Ki, "M": M, "world": world,
                         "us_local_gemm": round(us_local, 2), "us_nccl_gather": round(us_nccl, 2), "us_fused_gather": round(us_fused, 2), "us_fused_gather_torch_barrier": round(us_fused_torch_barrier, 2), "us_fused_gather_kernel_sync": (round(us_fused_kernel_sync, 2) if us_fused_kernel_sync is not None else None),
                         "fused_over_local": round(us_fused / us_local, 3), "TFLOPs_fused": round(flops / us_fused / 1e6, 1),
                         "gather_mode": "multicast" if world >= 4 else "peer stores",
                         "nvlink_bytes_sent_per_rank": sent if world >= 4 else sent * (world - 1),
                         "nvlink_bytes_received_per_rank": M * (No - (r1 - r0)) * 2,
                         "parity": {"gather_bit_identical": all_true(gather_ok and same_everywhere),
                                    "layer_bit_identical": all_true(bit_layer), "layer_rel_err": max_over_ranks(rel),
                                    "tolerance": 1e-2}})
            if not all_true(gather_ok and same_everywhere) or max_over_ranks(rel) >= 1e-2:
                raise RuntimeError(f"tensor-parallel parity failed: {rows[-1]}")
        del lin, linf, qf
        torch.cuda.empty_cache()
    return rows
