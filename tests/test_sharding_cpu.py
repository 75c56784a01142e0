"""CPU tests of the multi-GPU path's host logic: shard bounds,
the zero-communication property of row-sharded quantization (checked with the
oracle), and the all-gather of output columns over a world_size-2 gloo group."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_np as O  # noqa: E402
from quanta_b200.sharding import row_shard, shard_rows  # noqa: E402


@pytest.mark.parametrize("n,world,mult", [(4096, 8, 128), (14336, 8, 128), (1024, 4, 128), (1000, 3, 1), (28672, 8, 128),
                                          (200, 4, 128), (8192, 2, 128)])
def test_row_shard_partitions_rows(n, world, mult):
    bounds = [row_shard(n, world, r, mult) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
        assert a1 == b0 and a0 <= a1
    assert all(a % mult == 0 for a, b in bounds if b > a)
    sizes = [b - a for a, b in bounds if b > a]
    assert max(sizes) - min(sizes) <= mult or n % mult


def test_sharded_quantization_equals_slices_of_unsharded():
    """Blockwise-64 blocks run along in_features and never straddle a row: quantizing the row
    shards separately gives exactly the slices of the unsharded codes / scales (no collective)."""
    rng = np.random.default_rng(7)
    N, K, world = 512, 256, 4
    w = (rng.standard_normal((N, K)) * 0.02).astype(np.float32)
    pk, sc, zp = O.quantize4_block_pack(w, 64)
    for r in range(world):
        r0, r1 = row_shard(N, world, r, 128)
        pr, sr, zr = O.quantize4_block_pack(w[r0:r1], 64)
        assert np.array_equal(pr, pk.reshape(N, K // 2)[r0:r1].reshape(-1))
        assert np.array_equal(sr.view(np.uint32), sc.reshape(N, K // 64)[r0:r1].reshape(-1).view(np.uint32))
        assert np.array_equal(zr.view(np.uint32), zp.reshape(N, K // 64)[r0:r1].reshape(-1).view(np.uint32))


def _worker(rank, world, port, out_features, M, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from quanta_b200.sharding import gather_columns, row_shard as rs
        g = torch.Generator().manual_seed(0)
        x = torch.randn(M, 64, generator=g)
        w = torch.randn(out_features, 64, generator=g)
        full = x @ w.t()
        r0, r1 = rs(out_features, world, rank, 128)
        y_local = x @ w[r0:r1].t()                       # stands in for this rank's dequant-GEMM
        y = gather_columns(y_local.contiguous(), out_features, None, 128)
        results[rank] = bool(torch.equal(y, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("out_features", [256, 384, 200])
def test_gather_columns_world_size_2_gloo(out_features):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        results = mgr.dict()
        port = 29500 + (os.getpid() + out_features) % 2000
        procs = [ctx.Process(target=_worker, args=(r, 2, port, out_features, 5, results)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(results) == {0: True, 1: True}
