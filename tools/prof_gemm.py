"""One GEMM configuration a few times, for ncu.  args: N K M bits reps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16
N, K, M, bits, reps = [int(v) for v in (sys.argv[1:6] + ["14336", "4096", "16", "4", "4"][len(sys.argv) - 1:])]
torch.manual_seed(0)
w = torch.randn(N, K, device="cuda") * 0.02
q, s, z = Q.quantize_4bit(w, blocksize=64, packed=True) if bits == 4 else Q.quantize_8bit(w, blocksize=64)
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
for _ in range(reps):
    y = linear_wna16(x, q, s, z, None, bits=bits, blocksize=64, out_features=N)
torch.cuda.synchronize()
print("done", float(y.float().abs().max()))
