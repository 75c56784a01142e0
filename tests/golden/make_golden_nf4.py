"""Generate tests/golden/quanta_golden_nf4.npz by running the UNMODIFIED reference's NF4 path
(Quanta/functional/quantization.py:101-118, :59-61).  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_nf4.py

Per-tensor cases are the reference call itself; blockwise cases apply the reference function to
every block of B flat elements (how quanta_b200 defines ``blocksize=`` for nf4)."""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("QUANTA_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from Quanta.functional.quantization import quantize_4bit, dequantize_4bit  # noqa: E402

torch.set_num_threads(4)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quanta_golden_nf4.npz")
store, manifest = {}, []


def add(kind, x, block=None):
    name = f"nf4_{len(manifest):03d}"
    if block is None:
        idx, levels, am = quantize_4bit(x, quant_type="nf4")
        deq = dequantize_4bit(idx, levels, am, quant_type="nf4")
        store[f"{name}/absmax"] = am.numpy().reshape(())
    else:
        parts = [quantize_4bit(b, quant_type="nf4") for b in x.reshape(-1, block)]
        idx = torch.stack([p[0] for p in parts]).reshape(x.shape)
        levels = parts[0][1]
        store[f"{name}/absmax"] = torch.stack([p[2] for p in parts]).numpy()
        deq = torch.stack([dequantize_4bit(p[0], p[1], p[2], quant_type="nf4") for p in parts]).reshape(x.shape)
    store[f"{name}/x"] = x.numpy()
    store[f"{name}/idx"] = idx.numpy()
    store[f"{name}/deq"] = deq.numpy()
    store["levels"] = levels.numpy()
    manifest.append({"name": name, "kind": kind, "block": block})


def main():
    g = torch.Generator().manual_seed(4321)
    xs = [torch.tensor([-1.0, 0.0, 1.0, 2.0]), torch.randn(64, generator=g) * 0.02, torch.randn(7, 64, generator=g),
          torch.randn(16, 128, generator=g) * 0.02 + 0.5, torch.randn(33, 48, generator=g) * 3.0,
          torch.zeros(64), torch.full((64,), -2.0), torch.randn(4, 64, generator=g) * 1e-20,
          torch.randn(2, 64, generator=g) * 1e20]
    # values on and next to the decision boundaries between adjacent levels
    lv = quantize_4bit(torch.ones(1), quant_type="nf4")[1]
    mids = (lv[:-1] + lv[1:]) / 2
    edge = torch.cat([mids, torch.nextafter(mids, torch.tensor(2.0)), torch.nextafter(mids, torch.tensor(-2.0)), lv,
                      torch.tensor([1.0, -1.0, 0.0, -0.0, 1e-30])])
    xs.append(torch.cat([edge * 1.7, torch.zeros(128 - edge.numel() % 128)])[: (edge.numel() // 64) * 64 + 64])
    t = torch.randn(3, 64, generator=g)
    t[0, 5] = float("nan"); t[1, 7] = float("inf"); t[2] = 0.0
    xs.append(t)
    for x in xs:
        add("tensor", x)
        for B in (64, 128):
            if x.numel() % B == 0:
                add("block", x, B)
    store["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **store)
    print(f"wrote {OUT}: {len(manifest)} cases, torch {torch.__version__}")


if __name__ == "__main__":
    main()
