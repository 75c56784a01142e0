"""NF4 quantize / dequantize a few times on one Llama-2-7B matrix, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
x = torch.randn(11008, 4096, device="cuda") * 0.02
for _ in range(3):
    q, lv, am = Q.quantize_4bit(x, quant_type="nf4", blocksize=64, packed=True)
    d = Q.dequantize_4bit(q, lv, am, quant_type="nf4", blocksize=64, packed=True, shape=x.shape)
torch.cuda.synchronize()
print("done")
