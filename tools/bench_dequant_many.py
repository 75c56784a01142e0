"""quanta_dequantize_block_batch on one Llama-2-7B decoder layer (7 matrices, packed 4-bit block-64 -> fp32), CUDA-graph
replay over rotating copies: microseconds per launch and fraction of the measured copy peak.  QUANTA_B200_DQ_MULTI_CTAS_PER_SM
sets the persistent grid (8 = full occupancy is the default; 16 / 32 / 64 measure the same 177 us, 4: 256 us)."""
import sys, os, json, torch
sys.path.insert(0, os.getcwd())
import quanta_b200 as Q
layer = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
sets = []
for i in range(3):
    ws = [torch.randn(a, b, device="cuda") * 0.02 for (a, b) in layer]
    qs = Q.quantize_4bit_many(ws, blocksize=64, packed=True)
    sets.append(([t[0] for t in qs], [t[1] for t in qs], [t[2] for t in qs]))
    del ws
elems = sum(a * b for a, b in layer)
fn = lambda i: Q.dequantize_4bit_many(*sets[i % 3], blocksize=64, packed=True, shapes=layer)
for i in range(3): fn(i)
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side): fn(0)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    outs = [fn(i) for i in range(6)]
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 6
print(os.environ.get("QUANTA_B200_DQ_MULTI_CTAS_PER_SM"), round(us, 1), round(elems * 4.625 / us / 1e3 / 6459.6, 3))
