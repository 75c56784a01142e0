"""Worker of tests/test_gpu_tp.py: one rank of a tensor-parallel linear on its own GPU
(launched with torch.distributed.run).  Checks that both gather paths reproduce the
single-device result of the whole layer bit for bit."""
import os
import sys

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
g = torch.Generator().manual_seed(11)
for (N, K, M, bits) in [(1024, 512, 33, 4), (896, 1024, 7, 8), (2048, 256, 130, 4)]:
    w = torch.randn(N, K, generator=g) * 0.02
    b = torch.randn(N, generator=g) * 0.1
    x = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    qf = Q.quantize_4bit(w.to(dev), blocksize=64, packed=True) if bits == 4 else Q.quantize_8bit(w.to(dev), blocksize=64)
    for fused in (False, True, "peer", "multicast"):
        lin = TensorParallelLinear(K, N, bits=bits, bias=True, compute_dtype=torch.bfloat16, fused_gather=fused)
        r0, r1 = lin.rows
        lin.load_shard(w[r0:r1].to(dev), b[r0:r1].to(dev))
        for it in range(5):
            xi = (x * (1 + it)).contiguous()
            y = lin(xi).clone()
            ref = linear_wna16(xi, *qf, b.to(dev), bits=bits, blocksize=64, out_features=N)
            assert torch.equal(y, ref), f"rank {rank}: N={N} K={K} M={M} bits={bits} fused={fused} turn {it}"
# decode-sized batches on shapes whose K range is split over several CTAs: the ranks are synchronised INSIDE the kernel
# (quanta_gemm_wna16_scatter_sync).  Many back-to-back calls with changing inputs and no host synchronisation in
# between (both output buffers, epoch counting); the fused result must equal the all-gather of the same local GEMMs.
for (N, K, M, bits) in [(2048, 4096, 1, 4), (4096, 8192, 16, 4), (1024, 2048, 5, 8), (896, 11008, 9, 4)]:
    w = torch.randn(N, K, generator=g) * 0.02
    b = torch.randn(N, generator=g) * 0.1
    x = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    plain = TensorParallelLinear(K, N, bits=bits, bias=True, compute_dtype=torch.bfloat16, fused_gather=False)
    r0, r1 = plain.rows
    plain.load_shard(w[r0:r1].to(dev), b[r0:r1].to(dev))
    for fused in (True, "peer"):
        lin = TensorParallelLinear(K, N, bits=bits, bias=True, compute_dtype=torch.bfloat16, fused_gather=fused)
        lin.kernel_sync = True                                 # off by default (no faster): exercised here
        lin.qweight, lin.scale, lin.zero_point, lin.bias = plain.qweight, plain.scale, plain.zero_point, plain.bias
        xs = [(x * (1 + 0.25 * it)).contiguous() for it in range(24)]
        ys = [lin(xi).clone() for xi in xs]                    # 24 calls in a row
        torch.cuda.synchronize()
        assert lin._sym["epoch"] == 24, "the in-kernel synchronisation was not used"
        for it, xi in enumerate(xs):
            assert torch.equal(ys[it], plain(xi)), f"rank {rank}: N={N} K={K} M={M} bits={bits} fused={fused} call {it}"
# the same under CUDA-graph replay: the epoch of the in-kernel synchronisation lives in device memory and advances with
# every replayed launch (a host-side epoch would be frozen into the graph and the ranks would stop meeting)
N, K, M, bits = 2048, 4096, 4, 4
w = torch.randn(N, K, generator=g) * 0.02
xin = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
plain = TensorParallelLinear(K, N, bits=bits, bias=False, compute_dtype=torch.bfloat16, fused_gather=False)
r0, r1 = plain.rows
plain.load_shard(w[r0:r1].to(dev))
lin = TensorParallelLinear(K, N, bits=bits, bias=False, compute_dtype=torch.bfloat16, fused_gather=True)
lin.kernel_sync = True
lin.qweight, lin.scale, lin.zero_point = plain.qweight, plain.scale, plain.zero_point
xbuf = xin.clone()
for _ in range(2):
    lin(xbuf)                                              # symmetric buffers, workspaces exist before capture
torch.cuda.synchronize()
dist.barrier()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=side):
    outs = [lin(xbuf * (1.0 + 0.5 * j)).clone() for j in range(4)]
for rep in range(6):
    xbuf.copy_(xin * (1.0 + rep))
    graph.replay()
    torch.cuda.synchronize()
    for j in range(4):
        want = plain((xin * (1.0 + rep)) * (1.0 + 0.5 * j))
        assert torch.equal(outs[j], want), f"rank {rank}: graph replay {rep}, call {j}"
assert int(lin._sym["counter"].item()) == 2 + 4 * 6, int(lin._sym["counter"].item())
dist.barrier()
if rank == 0:
    print("TP_OK")
dist.destroy_process_group()
