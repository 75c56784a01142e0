"""int8 outlier matmul a few times at one shape, for ncu.  args: N K M reps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quanta_b200.nn import int8_outlier_matmul, rowwise_quantize_sym
N, K, M, reps = [int(v) for v in (sys.argv[1:5] + ["4096", "4096", "16", "3"][len(sys.argv) - 1:])]
torch.manual_seed(0)
qw, cw = rowwise_quantize_sym(torch.randn(N, K, device="cuda") * 0.02)
x = torch.randn(M, K, device="cuda")
x[:, [7, 513, 1024, 2049, 3071, 4000]] *= 20
x = x.to(torch.bfloat16)
for _ in range(reps):
    y = int8_outlier_matmul(x, qw, cw, threshold=6.0)
torch.cuda.synchronize()
print("done", float(y.float().abs().max()))
