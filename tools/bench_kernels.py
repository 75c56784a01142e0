"""Per-kernel roofline numbers: every hot-path entry point timed with CUDA
events on rotating inputs whose total footprint exceeds L2 (126 MB), after
warm-up.  Prints one JSON line per case: algorithmic bytes,
microseconds per call, GB/s and the fraction of the measured HBM peak.

    python tools/bench_kernels.py [--reps 20] [--out gpurun_out/kernels.jsonl]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200 import backends as QB

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--out", default=None)
args = ap.parse_args()

try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
dev = torch.device("cuda")
lines = []


def timeit(name, shape, alg_bytes, make_inputs, fn, copies):
    """GPU time per call under CUDA-graph replay (no host launch overhead) plus
    the eager per-call wall time through the Python API."""
    ins = [make_inputs(i) for i in range(copies)]
    for i in range(3):
        fn(ins[i % copies])
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for i in range(args.reps):
        fn(ins[i % copies])
    torch.cuda.synchronize()
    eager_us = (time.perf_counter() - t0) * 1e6 / args.reps
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(ins[0])                                   # per-stream workspace warm-up outside capture
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        outs = [fn(ins[i % copies]) for i in range(args.reps)]
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.reps
    gbs = alg_bytes / us / 1e3
    line = {"kernel": name, "shape": list(shape), "alg_bytes": alg_bytes, "us": round(us, 2), "GBps": round(gbs, 1),
            "frac_of_measured_hbm": round(gbs / PEAK, 3), "eager_us_per_call": round(eager_us, 1)}
    print(json.dumps(line), flush=True)
    lines.append(line)
    del ins, outs, g
    torch.cuda.empty_cache()


def randn(shape, i, scale=0.02):
    g = torch.Generator(device=dev).manual_seed(100 + i)
    return torch.randn(shape, device=dev, generator=g) * scale


for shape in [(4096, 4096), (11008, 4096)]:
    n = shape[0] * shape[1]
    copies = max(3, int(400e6 // (n * 4)) + 1)
    timeit("quantize_4bit block64 +pack (A3+P1)", shape, n * 4.625, lambda i: randn(shape, i),
           lambda x: Q.quantize_4bit(x, blocksize=64, packed=True), copies)
    timeit("quantize_4bit block64 unpacked (A3)", shape, n * 5.125, lambda i: randn(shape, i),
           lambda x: Q.quantize_4bit(x, blocksize=64), copies)
    timeit("quantize_8bit block64 (A3)", shape, n * 5.125, lambda i: randn(shape, i),
           lambda x: Q.quantize_8bit(x, blocksize=64), copies)
    timeit("quantize_8bit per-tensor (A1, two-pass)", shape, n * 5.0, lambda i: randn(shape, i, 1.0),
           lambda x: Q.quantize_8bit(x), copies)
    timeit("quantize_4bit per-tensor (A2, two-pass)", shape, n * 5.0, lambda i: randn(shape, i, 1.0),
           lambda x: Q.quantize_4bit(x), copies)
    timeit("quantize_8bit per_channel dim0 (A3, two-pass)", shape, n * 5.0, lambda i: randn(shape, i, 1.0),
           lambda x: Q.quantize_8bit(x, per_channel=True), copies)
    timeit("backends quantize_8bit sym per-tensor (B1)", shape, n * 5.0, lambda i: randn(shape, i, 1.0),
           lambda x: QB.quantize_8bit(x, False, True), copies)
    timeit("backends quantize_8bit asym per_channel (B1)", shape, n * 5.0, lambda i: randn(shape, i, 1.0),
           lambda x: QB.quantize_8bit(x, True, False), copies)

    def mk_q8(i):
        return Q.quantize_8bit(randn(shape, i, 1.0))
    timeit("dequantize_8bit per-tensor (A4)", shape, n * 5.0, mk_q8, lambda t: Q.dequantize_8bit(*t), copies)

    def mk_q8b(i):
        return Q.quantize_8bit(randn(shape, i), blocksize=64)
    timeit("dequantize_8bit block64 (A4)", shape, n * 5.125, mk_q8b, lambda t: Q.dequantize_8bit(*t, blocksize=64), copies)

    def mk_q4p(i):
        return Q.quantize_4bit(randn(shape, i), blocksize=64, packed=True)
    timeit("dequantize_4bit packed block64 -> fp32 (P2+A4)", shape, n * 4.625, mk_q4p,
           lambda t: Q.dequantize_4bit(*t, blocksize=64, packed=True, shape=shape), copies)
    timeit("dequantize_4bit packed block64 -> bf16", shape, n * 2.625, mk_q4p,
           lambda t: Q.dequantize_4bit(*t, blocksize=64, packed=True, shape=shape, out_dtype=torch.bfloat16), copies)

    def mk_b8(i):
        return QB.quantize_8bit(randn(shape, i, 1.0), False, True)
    timeit("backends dequantize_8bit sym (B2)", shape, n * 5.0, mk_b8, lambda t: QB.dequantize_8bit(*t), copies)

    def mk_codes(i):
        g = torch.Generator(device=dev).manual_seed(i)
        return torch.randint(0, 16, (n,), device=dev, dtype=torch.uint8, generator=g)
    timeit("pack_4bit_tensor (P1)", shape, n * 1.5, mk_codes, lambda c: Q.pack_4bit_tensor(c), max(copies, 12))
    timeit("unpack_4bit_tensor (P2)", shape, n * 1.5, lambda i: Q.pack_4bit_tensor(mk_codes(i))[0],
           lambda p: Q.unpack_4bit_tensor(p), max(copies, 20))
    xb = None

shape = (11008, 4096)
n = shape[0] * shape[1]
timeit("quantize_4bit block64 +pack, bf16 input", shape, n * 2.625, lambda i: randn(shape, i).to(torch.bfloat16),
       lambda x: Q.quantize_4bit(x, blocksize=64, packed=True), 6)

# paths that the tables above do not cover: per-channel dequantize, 16-bit inputs of the single-launch kernels
def mk_dim0(i):
    return Q.quantize_8bit(randn(shape, i), per_channel=True)


timeit("dequantize_8bit per_channel dim0 (A4)", shape, n * 5.0, mk_dim0, lambda t: Q.dequantize_8bit(*t), 3)
timeit("quantize_8bit per-tensor, bf16 input (A1)", shape, n * 3.0, lambda i: randn(shape, i).to(torch.bfloat16),
       lambda x: Q.quantize_8bit(x), 6)
timeit("quantize_8bit per_channel dim0, bf16 input (A3)", shape, n * 3.0, lambda i: randn(shape, i).to(torch.bfloat16),
       lambda x: Q.quantize_8bit(x, per_channel=True), 6)

# NF4 (row N1): per tensor (two passes) and blockwise, quantize and dequantize
timeit("quantize_4bit nf4 per-tensor (N1, two-pass)", shape, n * 5.0, lambda i: randn(shape, i),
       lambda x: Q.quantize_4bit(x, quant_type="nf4"), 3)
timeit("quantize_4bit nf4 block64 +pack (N1)", shape, n * (4 + 0.5 + 4 / 64), lambda i: randn(shape, i),
       lambda x: Q.quantize_4bit(x, quant_type="nf4", blocksize=64, packed=True), 3)


def mk_nf4(i):
    q, lv, am = Q.quantize_4bit(randn(shape, i), quant_type="nf4", blocksize=64, packed=True)
    return q, lv, am


timeit("dequantize_4bit nf4 packed block64 -> fp32 (N1)", shape, n * (0.5 + 4 + 4 / 64), mk_nf4,
       lambda t: Q.dequantize_4bit(*t, quant_type="nf4", blocksize=64, packed=True, shape=shape), 3)

# nf8 / fp4 / fp8 (row N4): one code byte per element
timeit("quantize_8bit nf8 per-tensor (N4, two-pass)", shape, n * 5.0, lambda i: randn(shape, i),
       lambda x: Q.quantize_8bit(x, quant_type="nf8"), 3)
timeit("quantize_8bit nf8 block64 (N4)", shape, n * (4 + 1 + 4 / 64), lambda i: randn(shape, i),
       lambda x: Q.quantize_8bit(x, quant_type="nf8", blocksize=64), 3)
timeit("quantize_8bit fp8 (N4)", shape, n * 5.0, lambda i: randn(shape, i),
       lambda x: Q.quantize_8bit(x, quant_type="fp8"), 3)
timeit("quantize_4bit fp4 (N4, one code per byte)", shape, n * 5.0, lambda i: randn(shape, i),
       lambda x: Q.quantize_4bit(x, quant_type="fp4"), 3)
timeit("dequantize_8bit fp8 -> fp32 (N4)", shape, n * 5.0, lambda i: Q.quantize_8bit(randn(shape, i), quant_type="fp8")[0],
       lambda q: Q.dequantize_8bit(q, None, 7, quant_type="fp8"), 3)
timeit("dequantize_8bit nf8 block64 -> fp32 (N4)", shape, n * (1 + 4 + 4 / 64),
       lambda i: Q.quantize_8bit(randn(shape, i), quant_type="nf8", blocksize=64),
       lambda t: Q.dequantize_8bit(*t, quant_type="nf8", blocksize=64), 3)

# batched blockwise quantize (one decoder layer's 7 matrices per call)
layer = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
nl = sum(a * b for a, b in layer)
timeit("quantize_4bit_many block64 +pack, one Llama-2-7B layer (7 matrices)", (nl,), nl * 4.625,
       lambda i: [randn(sh, 10 * i + j) for j, sh in enumerate(layer)],
       lambda ts: Q.quantize_4bit_many(ts, blocksize=64, packed=True), 2)

if args.out:
    with open(args.out, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
