"""The torch-eager restatement used as the CPU baseline (oracle/oracle_torch.py) must agree
bit for bit with the arithmetic oracle (oracle/oracle_np.py), which is pinned to the golden
vectors of the unmodified reference."""
import numpy as np
import torch

from oracle import oracle_np as O
from oracle import oracle_torch as OT


def test_blockwise_pack_matches_numpy_oracle():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(256, 512, generator=g) * 0.02
    x[0, :64] = 0.5                      # a degenerate block
    pk, sc, zp = OT.quantize4_block_pack(x, 64)
    po, so, zo = O.quantize4_block_pack(x.numpy(), 64)
    assert np.array_equal(pk.numpy(), po)
    assert np.array_equal(sc.numpy().view(np.uint32), so.view(np.uint32))
    assert np.array_equal(zp.numpy().view(np.uint32), zo.view(np.uint32))


def test_per_tensor_and_dequant_match_numpy_oracle():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(300, 70, generator=g)
    for bits in (8, 4):
        q, s, z = OT.quantize_linear(x, bits)
        qo, so, zo = O.quantize_affine(x.numpy(), bits, O.MODE_TENSOR)
        assert np.array_equal(q.numpy(), qo) and float(s) == float(so) and float(z) == float(zo)
        d = OT.dequantize_linear(q, s, z)
        assert np.array_equal(d.numpy().view(np.uint32), O.dequantize_affine(qo, so, zo).view(np.uint32))
    q, s, z = OT.quantize_linear(x, 8, per_channel=True)
    qo, so, zo = O.quantize_affine(x.numpy(), 8, O.MODE_DIM0)
    assert np.array_equal(q.numpy(), qo) and np.array_equal(s.numpy().reshape(-1), so.reshape(-1))
