"""Time one GEMM configuration (CUDA-graph of `reps` calls over rotating weights): N K M bits [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16
N, K, M, bits = [int(v) for v in sys.argv[1:5]]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
torch.manual_seed(0)
copies = max(4, int(400e6 // (N * K * bits // 8)) + 1)
ws = []
for i in range(copies):
    w = torch.randn(N, K, device="cuda") * 0.02
    ws.append(Q.quantize_4bit(w, blocksize=64, packed=True) if bits == 4 else Q.quantize_8bit(w, blocksize=64))
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
fn = lambda i: linear_wna16(x, *ws[i % copies], None, bits=bits, blocksize=64, out_features=N)
for i in range(3): fn(i)
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side): fn(0)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    outs = [fn(i) for i in range(reps)]
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"N={N} K={K} M={M} bits={bits} dbg={os.environ.get('QUANTA_B200_GEMM_DBG','0')} ctas={os.environ.get('QUANTA_B200_GEMM_CTAS','auto')}: {e0.elapsed_time(e1)*1e3/reps:.2f} us", flush=True)
