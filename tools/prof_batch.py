"""Small driver for ncu: the batched blockwise quantize on one Llama-2-7B decoder layer (7 matrices, one launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
torch.manual_seed(0)
layer = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
ts = [torch.randn(s, device="cuda") * 0.02 for s in layer]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = Q.quantize_4bit_many(ts, blocksize=64, packed=True)
torch.cuda.synchronize()
print("done", len(out))
