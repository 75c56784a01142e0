"""Diagnostic for the W8A16 llama-shape failure: separates GEMM error from dequantize error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16

torch.manual_seed(0)
for (N, K) in ((4096, 14336), (14336, 4096)):
    for bits in (8, 4):
        w = torch.randn(N, K, device="cuda") * 0.02
        x = torch.randn(256, K, device="cuda").to(torch.bfloat16)
        if bits == 4:
            q, s, z = Q.quantize_4bit(w, blocksize=64, packed=True)
            wd = Q.dequantize_4bit(q, s, z, blocksize=64, packed=True, shape=(N, K), out_dtype=torch.bfloat16)
            codes = Q.unpack_4bit_tensor(q)[: N * K].reshape(N, K)
        else:
            q, s, z = Q.quantize_8bit(w, blocksize=64)
            wd = Q.dequantize_8bit(q.reshape(N, K), s, z, blocksize=64, out_dtype=torch.bfloat16)
            codes = q.reshape(N, K)
        wt = (codes.float().reshape(-1, 64) * s[:, None] + z[:, None]).reshape(N, K).to(torch.bfloat16)
        dq_bad = int((wt != wd).sum())
        print(f"N={N} K={K} bits={bits}: dequant->bf16 mismatches vs torch: {dq_bad}", flush=True)
        for M in (1, 16, 64, 256):
            y = linear_wna16(x[:M], q, s, z, None, bits=bits, blocksize=64, out_features=N)
            torch.cuda.synchronize()
            ref = x[:M].float() @ wt.float().t()
            d = (y.float() - ref).abs()
            err = float(d.max() / ref.abs().max())
            tile_err = d.amax(dim=0).reshape(-1, 128).amax(dim=1) / ref.abs().max()
            bad = (tile_err > 1e-2).nonzero().flatten().tolist()
            print(f"   M={M}: rel_err={err:.3e} bad n-tiles={bad[:16]}{'...' if len(bad) > 16 else ''} ({len(bad)})", flush=True)
