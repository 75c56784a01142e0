"""Generate tests/golden/quanta_golden.npz by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not
exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every case stores the input tensor and the exact outputs of the reference
function named in ``fn`` (torch 2.11 CPU kernels).  Tests only ever read the
.npz; they never import the reference.

Case naming:  ``<row>_<idx>/<field>``;  a JSON manifest under key ``manifest``
lists (row, fn, kwargs) per case.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("QUANTA_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from Quanta.functional.quantization import (  # noqa: E402
    quantize_8bit, quantize_4bit, dequantize_8bit, dequantize_4bit)
from Quanta.backends.cpu.quantization import (  # noqa: E402
    quantize_8bit_cpu, quantize_4bit_cpu, dequantize_8bit_cpu, dequantize_4bit_cpu)
from Quanta.utils.utils import pack_4bit_tensor, unpack_4bit_tensor  # noqa: E402

torch.set_num_threads(4)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quanta_golden.npz")

store = {}
manifest = []


def npy(t):
    return t.detach().cpu().numpy().copy()


_inputs = {}


def add(row, fn, kwargs, **arrays):
    """Record one case; the input ``x`` is stored once per distinct tensor."""
    idx = len(manifest)
    name = f"{row}_{idx:03d}"
    entry = {"name": name, "row": row, "fn": fn, "kwargs": kwargs}
    for k, v in arrays.items():
        v = npy(v) if isinstance(v, torch.Tensor) else np.asarray(v)
        if k == "x":
            key = (v.shape, v.tobytes())
            if key not in _inputs:
                _inputs[key] = f"input_{len(_inputs):03d}"
                store[_inputs[key]] = v
            entry["x"] = _inputs[key]
        else:
            store[f"{name}/{k}"] = v
    manifest.append(entry)


def inputs():
    """Seeded inputs: LLM-like scaled randn, unscaled randn, shifted, tiny and
    large magnitudes, a few hand-made edge cases."""
    g = torch.Generator().manual_seed(1234)
    cases = []
    for shape, mul, add_ in [((4,), 1.0, 0.0), ((64,), 0.02, 0.0), ((7, 64), 0.02, 0.0),
                             ((33, 48), 1.0, 0.0), ((16, 128), 0.02, 0.5), ((8, 256), 3.0, -7.0),
                             ((5, 3, 64), 1.0, 0.0), ((96, 64), 1e-3, 0.0), ((12, 64), 1e4, 0.0),
                             ((2, 2, 4, 16), 0.5, 2.0)]:
        cases.append(torch.randn(shape, generator=g) * mul + add_)
    cases.append(torch.tensor([-1.0, 0.0, 1.0, 2.0]))                      # reference test vector
    cases.append(torch.tensor([[1.0, 2.0], [3.0, 4.0]]))                    # reference per_channel test
    cases.append(torch.zeros(64))
    cases.append(torch.full((64,), 2.0))
    cases.append(torch.full((128,), 1000.0))                                # scale == 0 path
    cases.append(torch.arange(128, dtype=torch.float32).reshape(2, 64))     # exact ties
    cases.append((torch.arange(256, dtype=torch.float32) * 0.5).reshape(4, 64))
    t = torch.randn(4, 64, generator=g)
    t[1] = 3.25                                                             # one constant block
    t[2, :32] = 0.0
    cases.append(t)
    cases.append(torch.randn(3, 64, generator=g).abs())                     # min >= 0
    return cases


def main():
    for x in inputs():
        n = x.numel()
        for bits, qf, df in ((8, quantize_8bit, dequantize_8bit), (4, quantize_4bit, dequantize_4bit)):
            # A1/A2 per tensor
            q, s, z = qf(x)
            add("A_tensor", f"functional.quantization.quantize_{bits}bit", {"bits": bits, "mode": "tensor"},
                x=x, q=q, scale=s, zp=z, deq=df(q, s, z))
            # A3 per_channel (dim 0)
            if x.dim() > 1:
                q, s, z = qf(x, per_channel=True)
                add("A_dim0", f"functional.quantization.quantize_{bits}bit(per_channel=True)",
                    {"bits": bits, "mode": "dim0"}, x=x, q=q, scale=s, zp=z, deq=df(q, s, z))
            # A3 blockwise = per_channel on x.reshape(-1,B).t()  
            for B in (32, 64, 128):
                if n % B == 0:
                    xt = x.reshape(-1, B).t()
                    q, s, z = qf(xt, per_channel=True)
                    deq = df(q, s, z)
                    add("A_block", f"functional.quantization.quantize_{bits}bit(x.reshape(-1,{B}).t(), per_channel=True)",
                        {"bits": bits, "mode": "block", "block": B}, x=x,
                        q=q.t().contiguous().reshape(x.shape), scale=s.reshape(-1), zp=z.reshape(-1),
                        deq=deq.t().contiguous().reshape(x.shape))
        # B1/B2
        for bits, qf, df in ((8, quantize_8bit_cpu, dequantize_8bit_cpu), (4, quantize_4bit_cpu, dequantize_4bit_cpu)):
            for pc in (False, True):
                if pc and x.dim() < 2:
                    continue
                for sym in (True, False):
                    q, s, z = qf(x, pc, sym)
                    add("B", f"backends.cpu.quantize_{bits}bit_cpu",
                        {"bits": bits, "per_channel": pc, "symmetric": sym},
                        x=x, q=q, scale=s, zp=z, deq=df(q, s, z))

    # B traps from SURVEY Appendix A.2
    for arr in ([0.0, 0.25, 0.5, 0.75, 1.0], [1.0, 1.00001], [1.0, 1.000009], [5.0, 5.0, 5.0]):
        x = torch.tensor(arr)
        for sym in (True, False):
            q, s, z = quantize_8bit_cpu(x, False, sym)
            add("B", "backends.cpu.quantize_8bit_cpu", {"bits": 8, "per_channel": False, "symmetric": sym},
                x=x, q=q, scale=s, zp=z, deq=dequantize_8bit_cpu(q, s, z))
    x = torch.tensor([[1.0, 2.0], [1.0, 3.0]])
    q, s, z = quantize_8bit_cpu(x, True, True)
    add("B", "backends.cpu.quantize_8bit_cpu", {"bits": 8, "per_channel": True, "symmetric": True},
        x=x, q=q, scale=s, zp=z, deq=dequantize_8bit_cpu(q, s, z))

    # P1/P2
    g = torch.Generator().manual_seed(99)
    for n in (2, 3, 16, 17, 64, 1001):
        qv = torch.randint(0, 16, (n,), generator=g, dtype=torch.uint8)
        packed, shape = pack_4bit_tensor(qv)
        add("P", "utils.pack_4bit_tensor", {"n": n}, q=qv, packed=packed, unpacked=unpack_4bit_tensor(packed))
    qv = torch.arange(16, dtype=torch.uint8)
    packed, _ = pack_4bit_tensor(qv)
    add("P", "utils.pack_4bit_tensor", {"n": 16}, q=qv, packed=packed, unpacked=unpack_4bit_tensor(packed))
    qv = torch.randint(0, 256, (64,), generator=g, dtype=torch.uint8)          # unmasked inputs > 15
    packed, _ = pack_4bit_tensor(qv)
    add("P", "utils.pack_4bit_tensor", {"n": 64, "unmasked": True}, q=qv, packed=packed,
        unpacked=unpack_4bit_tensor(packed))

    store["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **store)
    print(f"wrote {OUT}: {len(manifest)} cases, {os.path.getsize(OUT)/1024:.1f} KiB, torch {torch.__version__}")


if __name__ == "__main__":
    main()
