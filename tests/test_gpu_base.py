"""Convention C on the GPU (row N4): ``quanta_b200.functional.base.BaseQuantizer`` against the golden vectors of
the unmodified reference (tests/golden/quanta_golden_c.npz) and against the oracle on random inputs — codes,
scale / zero-point bit patterns and dequantized float32 bit patterns must all be identical."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quanta_golden_c.npz"))
CASES = json.loads(bytes(Z["manifest"]).decode())


def bits_equal(t, ref):
    """Same float32 bit patterns; NaNs only have to be NaNs on both sides (the sign / payload of a generated NaN
    differs between x86 and the GPU: 0xFFC00000 vs 0x7FFFFFFF)."""
    a = t.detach().cpu().numpy().astype(np.float32)
    b = np.asarray(ref, np.float32)
    if a.shape != b.shape:
        return False
    nan = np.isnan(a) & np.isnan(b)
    return bool(np.array_equal(np.where(nan, 0, a.view(np.uint32)), np.where(nan, 0, b.view(np.uint32))))


def test_golden_cases():
    from quanta_b200.functional.base import BaseQuantizer
    for c in CASES:
        x = torch.from_numpy(Z[c["name"] + "/x"]).cuda()
        bq = BaseQuantizer(c["bits"], c["symmetric"])
        q, s, z = bq.quantize(x, c["per_channel"])
        assert q.dtype == torch.uint8 and q.shape == x.shape and q.is_cuda
        assert np.array_equal(q.cpu().numpy(), Z[c["name"] + "/q"]), c
        assert bits_equal(s, Z[c["name"] + "/scale"]) and bits_equal(z, Z[c["name"] + "/zp"]), c
        d = bq.dequantize(q, s, z)
        assert bits_equal(d, Z[c["name"] + "/deq"]), c


@pytest.mark.parametrize("shape", [(1000,), (257, 4160), (4096, 1024), (3, 5, 64)])
@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("symmetric", [True, False])
def test_random_vs_oracle(shape, bits, symmetric):
    from quanta_b200.functional.base import BaseQuantizer
    g = torch.Generator().manual_seed(sum(shape) + bits)
    x = torch.randn(*shape, generator=g) * 0.3
    for per_channel in ((False, True) if len(shape) > 1 else (False,)):
        bq = BaseQuantizer(bits, symmetric)
        q, s, z = bq.quantize(x.cuda(), per_channel)
        qo, so, zo = O.base_quantize(x.numpy(), bits, per_channel, symmetric)
        assert np.array_equal(q.cpu().numpy(), qo)
        assert bits_equal(s, so) and bits_equal(z, zo)
        assert bits_equal(bq.dequantize(q, s, z), O.base_dequantize(qo, so, zo, bits, symmetric))


def test_api_mirrors_reference():
    from quanta_b200.functional.base import BaseQuantizer
    bq = BaseQuantizer()
    assert bq.num_bits == 8 and bq.symmetric is True and bq.max_val == 127
    assert BaseQuantizer(4, False).max_val == 15
    with pytest.raises(ValueError):
        bq.quantize(torch.randn(8, device="cuda"), per_channel=True)
    with pytest.raises(RuntimeError):
        bq.quantize(torch.randn(8))                       # CPU tensor: no fallback
