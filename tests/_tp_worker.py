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
dist.barrier()
if rank == 0:
    print("TP_OK")
dist.destroy_process_group()
