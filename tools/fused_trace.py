"""Phase timing of the single-launch per-tensor quantize (build with make EXTRA=-DQUANTA_FUSED_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
for rows in (512, 4096, 11008):
    xs = [torch.randn(rows, 4096, device="cuda") * 0.02 for _ in range(3)]
    torch.cuda.synchronize()
    for x in xs:
        print("rows", rows, flush=True)
        Q.quantize_8bit(x)
        torch.cuda.synchronize()
        Q.quantize_8bit(x, per_channel=True)
        torch.cuda.synchronize()
