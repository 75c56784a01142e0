"""Timeline of one gemm_small_kernel launch (debug build of the library: quanta_b200/csrc/build/variants/libT.so, built
with -DQUANTA_SMALL_TRACE by tools/build_variant.sh; see SM_TRACE in gemm_small.cu).  Prints, per CTA class, the clock64 stamps relative to the
CTA's entry and the globaltimer skew between CTAs."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import quanta_b200._lib as L
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(os.path.dirname(L.LIB_PATH), "csrc", "build", "variants", "libT.so"))
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--k", type=int, default=14336)
ap.add_argument("--m", type=int, default=1)
ap.add_argument("--bits", type=int, default=4)
ap.add_argument("--calls", type=int, default=6)
args = ap.parse_args()
L.LIB_PATH = args.lib
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16
N, K, M = args.n, args.k, args.m
ws = []
for i in range(max(args.calls, 4)):
    w = torch.randn(N, K, device="cuda") * 0.02
    ws.append(Q.quantize_4bit(w, blocksize=64, packed=True) if args.bits == 4 else Q.quantize_8bit(w, blocksize=64))
    del w
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
for i in range(3):
    linear_wna16(x, *ws[i % len(ws)], None, bits=args.bits, blocksize=64, out_features=N)
torch.cuda.synchronize()
# back-to-back calls from a graph (what bench_gemm times); the trace holds the LAST launch
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    outs = [linear_wna16(x, *ws[i % len(ws)], None, bits=args.bits, blocksize=64, out_features=N) for i in range(args.calls)]
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("us per call (graph of %d): %.2f" % (args.calls, e0.elapsed_time(e1) * 1e3 / args.calls))
lib = L.lib()
buf = np.zeros((148, 40), dtype=np.int64)
st = lib.quanta_debug_small_trace(C.c_void_p(buf.ctypes.data))
assert st == 0, st
g0 = buf[:, 38].min()
names = {32: "producer thread enters", 33: "weight barriers initialised", 34: "producer loop starts", 1: "producer first issue", 2: "producer dep-wait done", 3: "first stage landed", 4: "producer done", 5: "consumers enter", 24: "x warp 0 enters", 25: "x unit 0 staged", 26: "x unit 1 staged", 27: "cons: stage 0 full seen", 28: "cons: x 0 seen",
         6: "last segment loop end", 29: "epilogue: K quarters met", 30: "epilogue: result stored", 31: "epilogue: CTA synced", 7: "walk end", 36: "reducer: contributors seen", 37: "kernel end (tid 0)"}
np.set_printoptions(linewidth=200)
print("globaltimer entry skew (ns): min 0, median %d, max %d;  exit - first entry: median %d, max %d" % (
    np.median(buf[:, 38] - g0), (buf[:, 38] - g0).max(), np.median(buf[:, 39] - g0), (buf[:, 39] - g0).max()))
for slot, nm in names.items():
    v = buf[:, slot]
    print("%-28s cycles: min %6d  median %6d  max %6d" % (nm, v.min(), np.median(v), v.max()))
units = buf[:, 8:24]
for c in (0, 1, 73, 146, 147):
    u = units[c][units[c] > 0]
    print("cta %3d: entry +%d ns, units end at" % (c, buf[c, 38] - g0), u.tolist(), " deltas", np.diff(u).tolist())
    print("         first issue %d, landed %d, producer done %d, loop end %d, walk end %d, red seen %d, end %d" % tuple(buf[c, [1, 3, 4, 6, 7, 36, 37]]))
alld = np.concatenate([np.diff(units[c][units[c] > 0]) for c in range(148)])
print("unit-to-unit cycles over all CTAs: median %d, mean %d, p90 %d" % (np.median(alld), alld.mean(), np.percentile(alld, 90)))
