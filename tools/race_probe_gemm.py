"""Determinism probe of the GEMM paths: every repetition must reproduce the first result bit for bit
(stream-K reduces in a fixed order; int8 partial sums are exact)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16, linear_nf4a16, int8_outlier_matmul, rowwise_quantize_sym

torch.manual_seed(0)
for (N, K) in ((4096, 14336), (14336, 4096), (4096, 16384)):
    w = torch.randn(N, K, device="cuda") * 0.02
    q4 = Q.quantize_4bit(w, blocksize=64, packed=True)
    q8 = Q.quantize_8bit(w, blocksize=64)
    qn = Q.quantize_4bit(w, quant_type="nf4", blocksize=64, packed=True)
    qi = rowwise_quantize_sym(w)
    for M in (1, 16, 64, 256, 300):
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        xo = x.clone(); xo[:, [7, 513, 1024]] *= 20
        fns = {"w4": lambda: linear_wna16(x, *q4, None, bits=4, blocksize=64, out_features=N),
               "w8": lambda: linear_wna16(x, *q8, None, bits=8, blocksize=64, out_features=N),
               "nf4": lambda: linear_nf4a16(x, qn[0], qn[2], None, blocksize=64, out_features=N),
               "int8+outliers": lambda: int8_outlier_matmul(xo, qi[0], qi[1], threshold=6.0)}
        for name, fn in fns.items():
            ref = fn().clone()
            bad = sum(0 if torch.equal(fn(), ref) else 1 for _ in range(100))
            if bad:
                print(f"N={N} K={K} M={M} {name}: {bad} of 100 runs differ", flush=True)
    print(f"N={N} K={K} done", flush=True)
