"""Config 3: fused W4A16 / W8A16 dequant-GEMM on Llama-3-8B linear shapes,
M = 1..256.  CUDA-graph replay over rotating weight copies (total footprint
larger than L2, so weights stream from HBM on every call), CUDA events.
Reports per point: microseconds, TFLOP/s, algorithmic GB/s and the fraction
of the binding roof (HBM below the ridge, tensor above) from MEASURED_PEAKS.json."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16, linear_nf4a16

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=24)
ap.add_argument("--out", default=None)
ap.add_argument("--ms", default="1,8,16,32,64,128,256")
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--formats", default="4,8", help="comma list of 4, 8, nf4")
ap.add_argument("--shapes", default="4096x14336,14336x4096")
args = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    pk = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    HBM, TF, TFS = pk["hbm_gbs"], pk["bf16_tflops"], pk["bf16_tflops_sustained"]
except Exception:
    HBM, TF, TFS = 6650.0, 1590.0, 1400.0
dt = torch.bfloat16 if args.dtype == "bf16" else torch.float16
lines = []
for (N, K) in [tuple(int(v) for v in sh.split('x')) for sh in args.shapes.split(',')]:
    for fmt in args.formats.split(","):
        nf4 = fmt == "nf4"
        bits = 4 if nf4 else int(fmt)
        wbytes = N * K * bits // 8 + (N * K // 64) * (4 if nf4 else 8)
        copies = max(4, int(400e6 // wbytes) + 1)
        ws = []
        for i in range(copies):
            w = torch.randn(N, K, device="cuda") * 0.02
            if nf4:
                q_, _, am_ = Q.quantize_4bit(w, quant_type="nf4", blocksize=64, packed=True)
                ws.append((q_, am_))
            else:
                ws.append(Q.quantize_4bit(w, blocksize=64, packed=True) if bits == 4 else Q.quantize_8bit(w, blocksize=64))
            del w
        for M in [int(v) for v in args.ms.split(",")]:
            x = torch.randn(M, K, device="cuda").to(dt)
            if nf4:
                fn = lambda i: linear_nf4a16(x, *ws[i % copies], None, blocksize=64, out_features=N)
            else:
                fn = lambda i: linear_wna16(x, *ws[i % copies], None, bits=bits, blocksize=64, out_features=N)
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(0)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                outs = [fn(i) for i in range(args.reps)]
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / args.reps
            flops = 2.0 * M * N * K
            abytes = wbytes + 2 * M * K + 2 * M * N
            t_hbm, t_tc = abytes / HBM / 1e3, flops / TFS / 1e6          # microseconds
            bound = "hbm" if t_hbm >= t_tc else "tensor"
            line = {"op": "NF4A16" if nf4 else f"W{bits}A16", "N": N, "K": K, "M": M, "dtype": args.dtype, "us": round(us, 2),
                    "TFLOPs": round(flops / us / 1e6, 1), "GBps": round(abytes / us / 1e3, 1), "bound": bound,
                    "frac_of_roof": round(max(t_hbm, t_tc) / us, 3), "t_hbm_us": round(t_hbm, 2), "t_tensor_us": round(t_tc, 2)}
            print(json.dumps(line), flush=True)
            lines.append(line)
            del g, outs
        del ws
        torch.cuda.empty_cache()
if args.out:
    with open(args.out, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
