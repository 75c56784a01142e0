"""Does the second pass of the two-pass (per-tensor) quantize hit L2?  Times quantize_8bit per tensor
for growing sizes on rotating inputs (footprint > L2) under CUDA-graph replay and prints the
moved-bytes bandwidth under both assumptions (second read from DRAM: 9 B/elem; from L2: 5 B/elem)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q

dev = torch.device("cuda")
reps = 20
for rows in (512, 1024, 2048, 3072, 4096, 6144, 8192):
    shape = (rows, 4096)
    n = rows * 4096
    copies = max(3, int(600e6 // (n * 4)) + 1)
    ins = [torch.randn(shape, device=dev) * 0.02 for _ in range(copies)]
    for mode in ("tensor", "block64"):
        fn = (lambda x: Q.quantize_8bit(x)) if mode == "tensor" else (lambda x: Q.quantize_8bit(x, blocksize=64))
        for i in range(3):
            fn(ins[i % copies])
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(ins[0])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            outs = [fn(ins[i % copies]) for i in range(reps)]
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(json.dumps({"mode": mode, "MiB": n * 4 / 2**20, "us": round(us, 2), "GBps_5B": round(n * 5 / us / 1e3, 1),
                          "GBps_9B": round(n * 9 / us / 1e3, 1)}), flush=True)
        del g, outs
    del ins
    torch.cuda.empty_cache()
