"""NF4 kernels only (subset of tools/bench_kernels.py)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
dev = torch.device("cuda")
reps = 20
shape = (11008, 4096)
n = shape[0] * shape[1]
copies = 3
ins = [torch.randn(shape, device=dev) * 0.02 for _ in range(copies)]
qs = [Q.quantize_4bit(x, quant_type="nf4", blocksize=64, packed=True) for x in ins]


def run(name, fn, args, alg):
    for i in range(3):
        fn(args[i % copies])
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        outs = [fn(args[i % copies]) for i in range(reps)]
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(json.dumps({"kernel": name, "us": round(us, 2), "GBps": round(alg / us / 1e3, 1), "frac": round(alg / us / 1e3 / 6459.6, 3)}), flush=True)


run("nf4 quantize block64+pack", lambda x: Q.quantize_4bit(x, quant_type="nf4", blocksize=64, packed=True), ins, n * 4.5625)
run("nf4 quantize per-tensor", lambda x: Q.quantize_4bit(x, quant_type="nf4"), ins, n * 5.0)
run("nf4 dequantize packed block64 -> fp32", lambda t: Q.dequantize_4bit(t[0], t[1], t[2], quant_type="nf4", blocksize=64, packed=True, shape=shape), qs, n * 4.5625)
