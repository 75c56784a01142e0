"""Small driver for ncu: runs each hot kernel a few times on representative
shapes (one Llama-2-7B matrix), so a profile costs seconds instead of the whole
25.9 GB bench.  Usage: python tools/prof_kernels.py [what] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
x = torch.randn(11008, 4096, device="cuda") * 0.02
x2 = torch.randn(4096, 4096, device="cuda")

for _ in range(reps):
    if what in ("all", "block4"):
        pk, s, z = Q.quantize_4bit(x, blocksize=64, packed=True)
    if what in ("all", "block8"):
        q8, s8, z8 = Q.quantize_8bit(x, blocksize=64)
    if what in ("all", "tensor8"):
        q, sc, zp = Q.quantize_8bit(x2)
        d = Q.dequantize_8bit(q, sc, zp)
    if what in ("all", "dim0"):
        q, sc, zp = Q.quantize_8bit(x2, per_channel=True)
    if what in ("all", "dequant4"):
        pk, s, z = Q.quantize_4bit(x, blocksize=64, packed=True)
        d = Q.dequantize_4bit(pk, s, z, blocksize=64, packed=True, shape=x.shape)
    if what in ("all", "pack"):
        c = torch.randint(0, 16, (11008 * 4096,), device="cuda", dtype=torch.uint8)
        p, _ = Q.pack_4bit_tensor(c)
        u = Q.unpack_4bit_tensor(p)
torch.cuda.synchronize()
print("done", what)
