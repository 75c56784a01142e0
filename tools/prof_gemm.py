"""One dequant-GEMM shape a few times in a row (for ncu):  python tools/prof_gemm.py N K M bits [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16

N, K, M, bits = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 6
ws = []
for i in range(4):
    w = torch.randn(N, K, device="cuda") * 0.02
    ws.append(Q.quantize_4bit(w, blocksize=64, packed=True) if bits == 4 else Q.quantize_8bit(w, blocksize=64))
    del w
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
torch.cuda.synchronize()
for i in range(reps):
    y = linear_wna16(x, *ws[i % 4], None, bits=bits, blocksize=64, out_features=N)
torch.cuda.synchronize()
print("ok", float(y.float().abs().max()))
