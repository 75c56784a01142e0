"""2+ GPU probe: fused gather through the NVSwitch multicast mapping (one store per tile instead of one per peer).
torchrun --nproc-per-node N tools/tp_multicast_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import quanta_b200 as Q
from quanta_b200.nn import linear_wna16
from quanta_b200.sharding import TensorParallelLinear

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(3)
N, K = 8192, 8192
w = torch.randn(N, K, generator=g) * 0.02
qf = Q.quantize_4bit(w.to(dev), blocksize=64, packed=True)
layers = {}
for mode in ("peer", True):
    lin = TensorParallelLinear(K, N, bits=4, bias=False, compute_dtype=torch.bfloat16, fused_gather=mode)
    r0, r1 = lin.rows
    lin.load_shard(w[r0:r1].to(dev))
    layers[mode] = lin
for M in (16, 256):
    x = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    ref = linear_wna16(x, *qf, None, bits=4, blocksize=64, out_features=N)
    first = None
    for mode, lin in layers.items():
        y = lin(x).clone()
        used_mc = bool(lin._sym["mc"])
        # the per-peer and the multicast gather run the same local GEMM: their results must be identical; against
        # the single-device layer only the stream-K split (hence the fp32 summation order) may differ
        if first is None:
            first = y
        ok = bool(torch.equal(y, first)) and float((y.float() - ref.float()).abs().max() / ref.float().abs().max()) < 1e-2
        for _ in range(3):
            lin(x)
        torch.cuda.synchronize(); dist.barrier()
        gr = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            lin(x)
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize(); dist.barrier()
        with torch.cuda.graph(gr):
            for _ in range(20):
                lin(x)
        gr.replay(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"M={M} mode={mode} multicast_used={used_mc} bit_identical={ok} us={t.item():.2f}", flush=True)
dist.barrier()
dist.destroy_process_group()
