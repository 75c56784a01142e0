"""Config 4: LLM.int8()-style outlier-split matmul on OPT-6.7B layer shapes, tokens in
{1, 16, 256, 2048}; activations randn with 6 fixed feature columns multiplied by 20 so that threshold = 6.0
splits non-trivially.  CUDA-graph replay over rotating weight copies (> L2), CUDA events.  Reports
microseconds, TFLOP/s (2*M*N*K) and the fraction of the binding roof: HBM (int8 weights + activations +
output) below the ridge, else an INT8 tensor peak MEASURED in the same run: cuBLASLt's int8 GEMM
(torch._int_mm, 8192^3, best of 10, CUDA events) — the library figure a kind::i8 kernel is judged against, as
the bf16 figures are judged against cuBLAS bf16 (round 1 divided by the bf16 peak, which overstated these
fractions about twofold)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quanta_b200.nn import int8_outlier_matmul, rowwise_quantize_sym

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--out", default=None)
args = ap.parse_args()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    pk = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    HBM, TFS = pk["hbm_gbs"], pk["bf16_tflops_sustained"]
except Exception:
    HBM, TFS = 6650.0, 1400.0
OUTL = [7, 513, 1024, 2049, 3071, 4000]
lines = []


def measure_int8_peak():
    a = torch.randint(-127, 127, (8192, 8192), device="cuda", dtype=torch.int8)
    b = torch.randint(-127, 127, (8192, 8192), device="cuda", dtype=torch.int8).t()        # column-major B, as cuBLASLt wants
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12


try:
    INT8_PEAK = measure_int8_peak()
    peak_line = {"int8_tensor_peak_TOPs": round(INT8_PEAK, 1), "how": "torch._int_mm (cuBLASLt int8) 8192^3, best of 10, burst"}
except Exception as ex:                                            # noqa: BLE001
    INT8_PEAK = 2.0 * TFS
    peak_line = {"int8_tensor_peak_TOPs": round(INT8_PEAK, 1), "how": f"fallback 2 x sustained bf16 peak ({type(ex).__name__})"}
print(json.dumps(peak_line), flush=True)
lines.append(peak_line)
for (N, K) in ((4096, 4096), (16384, 4096), (4096, 16384)):
    copies = max(3, int(300e6 // (N * K)) + 1)
    ws = []
    for i in range(copies):
        w = torch.randn(N, K, device="cuda") * 0.02
        ws.append(rowwise_quantize_sym(w))
    for M in (1, 16, 256, 2048):
        x = torch.randn(M, K, device="cuda")
        x[:, OUTL] *= 20.0
        x = x.to(torch.bfloat16)
        fn = lambda i: int8_outlier_matmul(x, ws[i % copies][0], ws[i % copies][1], threshold=6.0)
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(0)
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            outs = [fn(i) for i in range(args.reps)]
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        flops = 2.0 * M * N * K
        nbytes = N * K + 4 * N + 2 * M * K + 2 * M * N
        t_hbm, t_tensor = nbytes / HBM / 1e3, flops / INT8_PEAK / 1e6
        bound = "hbm" if t_hbm >= t_tensor else "tensor(measured int8 peak)"
        line = {"op": "int8_outlier_matmul", "N": N, "K": K, "M": M, "us": round(us, 2), "TFLOPs": round(flops / us / 1e6, 1),
                "GBps": round(nbytes / us / 1e3, 1), "bound": bound, "frac_of_roof": round(max(t_hbm, t_tensor) / us, 3),
                "launches_per_call": 4}
        print(json.dumps(line), flush=True)
        lines.append(line)
        del g, outs
    del ws
    torch.cuda.empty_cache()
if args.out:
    with open(args.out, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
