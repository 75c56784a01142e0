"""One GEMM case: python tools/gemm_case.py N K M bits [dtype] — prints the relative error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16

N, K, M, bits = [int(v) for v in sys.argv[1:5]]
dt = torch.float16 if (len(sys.argv) > 5 and sys.argv[5] == "fp16") else torch.bfloat16
torch.manual_seed(0)
w = torch.randn(N, K, device="cuda") * 0.02
x = torch.randn(M, K, device="cuda").to(dt)
if bits == 4:
    q, s, z = Q.quantize_4bit(w, blocksize=64, packed=True)
    wd = Q.dequantize_4bit(q, s, z, blocksize=64, packed=True, shape=(N, K), out_dtype=dt)
else:
    q, s, z = Q.quantize_8bit(w, blocksize=64)
    wd = Q.dequantize_8bit(q, s, z, blocksize=64, out_dtype=dt)
ref = x.float() @ wd.float().t()
torch.cuda.synchronize()
for rep in range(2):
    y = linear_wna16(x, q, s, z, None, bits=bits, blocksize=64, out_features=N)
    torch.cuda.synchronize()
    err = float((y.float() - ref).abs().max() / ref.abs().max())
    print(f"N={N} K={K} M={M} bits={bits} {dt} G={os.environ.get('QUANTA_B200_GEMM_CTAS','auto')}: rel_err={err:.3e} {'OK' if err < 1e-2 else 'BAD'}", flush=True)
