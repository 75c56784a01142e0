"""Generate tests/golden/quanta_golden_c.npz by running the UNMODIFIED reference's BaseQuantizer
(Quanta/functional/base.py:5-72, convention C of SURVEY Appendix A.3).  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_c.py

Every case stores the input and the exact (q, scale, zero_point, dequantized) of
``BaseQuantizer(num_bits, symmetric).quantize(x, per_channel)`` / ``.dequantize(...)``."""
import json
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("QUANTA_REFERENCE", "/root/reference"))
from Quanta.functional.base import BaseQuantizer  # noqa: E402

torch.set_num_threads(4)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quanta_golden_c.npz")
store, manifest = {}, []
g = torch.Generator().manual_seed(20261018)


def add(x, bits, symmetric, per_channel):
    bq = BaseQuantizer(bits, symmetric)
    q, s, z = bq.quantize(x, per_channel)
    d = bq.dequantize(q, s, z)
    name = f"C_{len(manifest):03d}"
    manifest.append({"name": name, "bits": bits, "symmetric": symmetric, "per_channel": per_channel})
    for k, v in (("x", x), ("q", q), ("scale", s), ("zp", z), ("deq", d)):
        store[f"{name}/{k}"] = v.detach().cpu().numpy().copy()


inputs = [
    torch.tensor([-1.0, -0.5, 0.0, 0.5, 1.0]),
    torch.arange(1, 10, dtype=torch.float32).reshape(3, 3),          # Appendix B (per_channel)
    torch.tensor([-1.0, 0.0, 1.0, 2.0]),
    torch.randn(37, 29, generator=g),
    torch.randn(64, 96, generator=g) * 0.02,
    torch.randn(5, 7, 9, generator=g) * 3.0,
    torch.rand(130, 33, generator=g),                                 # all positive: asymmetric zp = min > 0
    torch.full((4, 6), 2.0),                                          # allclose(min, max): scale 1, zp = min, codes computed
    torch.zeros(3, 5),
    torch.full((8,), -3.25),
    torch.tensor([[1.0, 2.0], [1.0, 3.0]]),                           # one degenerate channel only
    torch.tensor([[0.0, -0.0, 1e-30, -1e-30, 3.4e38, -3.4e38]]),
    torch.randn(16, 8, generator=g) * 1e-6,
    torch.randn(16, 8, generator=g) * 1e6,
]
for x in inputs:
    for bits in (8, 4):
        for symmetric in (True, False):
            add(x, bits, symmetric, False)
            if x.dim() > 1:
                add(x, bits, symmetric, True)
store["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
np.savez_compressed(OUT, **store)
print(f"wrote {OUT}: {len(manifest)} cases")
