"""Config 5: row-sharded quantization + tensor-parallel dequant-GEMM of
Llama-3-70B-shaped weights on N GPUs of one box.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/bench_tp.py [--out file.jsonl]

Every rank holds rows row_shard(out, N, rank, 128) of each weight.  Reports (rank 0, max over
ranks, CUDA events): quantize GB/s aggregate (no collective on that path) and the TP linear
(local fused kernel + one NCCL all-gather, and the kernel with the gather fused into its
epilogue over symmetric memory) in microseconds per call for M in {1,16,64,256}.
A small case is first checked against a single-device computation of the whole layer."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import quanta_b200 as Q
from quanta_b200.sharding import TensorParallelLinear, row_shard

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lines = []


def emit(d):
    if rank == 0:
        print(json.dumps(d), flush=True)
        lines.append(d)


def max_over_ranks(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- correctness: TP result == single-device result of the whole layer ----
N, K, M = 1024, 512, 33
g = torch.Generator().manual_seed(1)
w = torch.randn(N, K, generator=g) * 0.02
b = torch.randn(N, generator=g) * 0.1
x = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
for bits in (4, 8):
    lin = TensorParallelLinear(K, N, bits=bits, bias=True, compute_dtype=torch.bfloat16)
    r0, r1 = lin.rows
    lin.load_shard(w[r0:r1].to(dev), b[r0:r1].to(dev))
    y = lin(x)
    qf = Q.quantize_4bit(w.to(dev), blocksize=64, packed=True) if bits == 4 else Q.quantize_8bit(w.to(dev), blocksize=64)
    from quanta_b200.nn import linear_wna16
    ref = linear_wna16(x, *qf, b.to(dev), bits=bits, blocksize=64, out_features=N)
    same = bool(torch.equal(y, ref))
    err = float((y.float() - ref.float()).abs().max() / ref.float().abs().max())
    emit({"check": f"tp_linear W{bits}A16 == single device", "world": world, "bit_identical": same, "rel_err": err})
    assert err < 1e-2
    if world > 1:
        linf = TensorParallelLinear(K, N, bits=bits, bias=True, compute_dtype=torch.bfloat16, fused_gather=True)
        linf.load_shard(w[r0:r1].to(dev), b[r0:r1].to(dev))
        ok = True
        for it in range(6):                                    # both buffers, several turns
            xi = (x * (1 + it)).contiguous()
            yf = linf(xi).clone()
            refi = linear_wna16(xi, *qf, b.to(dev), bits=bits, blocksize=64, out_features=N)
            ok = ok and bool(torch.equal(yf, refi))
        emit({"check": f"tp_linear W{bits}A16 fused gather == single device", "world": world, "bit_identical": ok})
        assert ok
        del linf

# ---- throughput on Llama-3-70B shapes ----
SHAPES = [(8192, 8192), (1024, 8192), (28672, 8192), (8192, 28672)]
for (No, Ki) in SHAPES:
    r0, r1 = row_shard(No, world, rank, 128)
    rows = r1 - r0
    wl = torch.empty(rows, Ki, device=dev).normal_(0.0, 0.02)
    # row-sharded 4-bit block-64 quantize+pack: no communication
    for _ in range(3):
        q4 = Q.quantize_4bit(wl, blocksize=64, packed=True)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        q4 = Q.quantize_4bit(wl, blocksize=64, packed=True)
    e1.record(); torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1) / 10)
    emit({"op": "quantize_4bit block64+pack, row-sharded", "shape": [No, Ki], "world": world,
          "us": round(ms * 1e3, 2), "aggregate_GBps": round(No * Ki * 4.625 / (ms * 1e-3) / 1e9, 1)})
    for bits in (4, 8):
        lin = TensorParallelLinear(Ki, No, bits=bits, bias=False, compute_dtype=torch.bfloat16)
        lin.load_shard(wl)
        linf = linp = None
        if world > 1:
            linf = TensorParallelLinear(Ki, No, bits=bits, bias=False, compute_dtype=torch.bfloat16, fused_gather=True)
            linf.qweight, linf.scale, linf.zero_point = lin.qweight, lin.scale, lin.zero_point
            linp = TensorParallelLinear(Ki, No, bits=bits, bias=False, compute_dtype=torch.bfloat16, fused_gather="peer")
            linp.qweight, linp.scale, linp.zero_point = lin.qweight, lin.scale, lin.zero_point
        for Mb in (1, 16, 64, 256):
            xb = torch.randn(Mb, Ki, device=dev).to(torch.bfloat16)
            for _ in range(3):
                lin(xb)
            torch.cuda.synchronize(); dist.barrier()
            e0.record()
            for _ in range(args.reps):
                yl = lin.local_matmul(xb)
            e1.record(); torch.cuda.synchronize()
            us_local = max_over_ranks(e0.elapsed_time(e1) * 1e3 / args.reps)
            dist.barrier()
            e0.record()
            for _ in range(args.reps):
                y = lin(xb)
            e1.record(); torch.cuda.synchronize()
            us = max_over_ranks(e0.elapsed_time(e1) * 1e3 / args.reps)
            rec = {"op": f"tp_linear W{bits}A16", "N": No, "K": Ki, "M": Mb, "world": world, "us": round(us, 2),
                   "us_local_gemm": round(us_local, 2), "TFLOPs": round(2.0 * Mb * No * Ki / us / 1e6, 1)}
            if linf is not None:
                for _ in range(3):
                    linf(xb)
                torch.cuda.synchronize(); dist.barrier()
                e0.record()
                for _ in range(args.reps):
                    y = linf(xb)
                e1.record(); torch.cuda.synchronize()
                rec["us_fused"] = round(max_over_ranks(e0.elapsed_time(e1) * 1e3 / args.reps), 2)
                # the same 20 calls replayed from a CUDA graph: device time without the host's launch path
                for name, layer in (("us_graph", lin), ("us_fused_graph", linf), ("us_fused_peer_graph", linp)):
                    try:
                        side = torch.cuda.Stream()
                        side.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(side):
                            layer(xb)
                        torch.cuda.current_stream().wait_stream(side)
                        torch.cuda.synchronize(); dist.barrier()
                        gr = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gr):
                            for _ in range(args.reps):
                                y = layer(xb)
                        gr.replay(); torch.cuda.synchronize(); dist.barrier()
                        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
                        rec[name] = round(max_over_ranks(e0.elapsed_time(e1) * 1e3 / args.reps), 2)
                        del gr
                    except Exception as ex:                     # noqa: BLE001
                        rec[name] = f"failed: {type(ex).__name__}: {str(ex)[:80]}"
                        torch.cuda.synchronize()
            emit(rec)
        del lin, linf
        linp = None
    del wl, q4
    torch.cuda.empty_cache()
if rank == 0 and args.out:
    with open(args.out, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
dist.destroy_process_group()
