"""Quick GEMM diagnostic: small cases first, prints relative error per case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16

torch.manual_seed(0)
cases = [(128, 64, 16, 8), (128, 64, 16, 4), (128, 256, 16, 4), (256, 512, 5, 4), (384, 1024, 33, 8), (1024, 4096, 64, 4),
         (512, 2048, 256, 4), (4096, 14336, 16, 4), (14336, 4096, 256, 4), (4096, 14336, 256, 8)]
for N, K, M, bits in cases:
    for dt in (torch.bfloat16, torch.float16):
        w = torch.randn(N, K, device="cuda") * 0.02
        x = torch.randn(M, K, device="cuda").to(dt)
        if bits == 4:
            q, s, z = Q.quantize_4bit(w, blocksize=64, packed=True)
            wd = Q.dequantize_4bit(q, s, z, blocksize=64, packed=True, shape=(N, K), out_dtype=dt)
        else:
            q, s, z = Q.quantize_8bit(w, blocksize=64)
            wd = Q.dequantize_8bit(q, s, z, blocksize=64, out_dtype=dt)
        y = linear_wna16(x, q, s, z, None, bits=bits, blocksize=64, out_features=N)
        torch.cuda.synchronize()
        ref = x.float() @ wd.float().t()
        err = float((y.float() - ref).abs().max() / ref.abs().max())
        print(f"N={N} K={K} M={M} bits={bits} {dt}: rel_err={err:.3e} {'OK' if err < 1e-2 else 'BAD'}", flush=True)
