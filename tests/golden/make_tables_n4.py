"""Derive the exact decision tables of the reference's nf8 / fp4 / fp8 paths and their golden vectors by
running the UNMODIFIED reference (Quanta/functional/quantization.py:120-183, :39-49, :62-69) in the
build container (torch CPU):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_tables_n4.py

Writes  tests/golden/quanta_tables_n4.npz   nf8 levels (256), nf8 thresholds (255), fp4 / fp8 exponent thresholds
        tests/golden/quanta_golden_n4.npz   inputs and reference outputs (codes, dequantized values)

A threshold is the smallest float for which the reference's decision reaches the next code; the decisions
are monotone, which the script verifies on a neighbourhood of every threshold and on random samples."""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("QUANTA_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from Quanta.functional.quantization import quantize_4bit, quantize_8bit, dequantize_4bit, dequantize_8bit  # noqa: E402

torch.set_num_threads(4)
HERE = os.path.dirname(os.path.abspath(__file__))


def f2i(x):
    """float32 -> int64 key with the same order (negative floats mirrored)."""
    b = x.view(torch.int32).to(torch.int64)
    return torch.where(b < 0, -(b & 0x7FFFFFFF), b)


def i2f(k):
    b = torch.where(k < 0, (-k) | 0x80000000, k)
    return (b & 0xFFFFFFFF).to(torch.int64).to(torch.uint32).view(torch.float32) if hasattr(torch, "uint32") else None


def i2f_np(k):
    k = k.numpy().astype(np.int64)
    b = np.where(k < 0, (-k) | 0x80000000, k).astype(np.uint32)
    return torch.from_numpy(b.view(np.float32).copy())


def bisect(decide, lo, hi, target):
    """Smallest float v in (lo, hi] with decide(v) >= target, for vectors of lo / hi / target (decide monotone)."""
    lo_k, hi_k = f2i(lo), f2i(hi)
    assert bool((decide(lo) < target).all()) and bool((decide(hi) >= target).all())
    while bool((hi_k - lo_k > 1).any()):
        mid = (lo_k + hi_k) // 2
        ok = decide(i2f_np(mid)) >= target
        hi_k = torch.where(ok, mid, hi_k)
        lo_k = torch.where(ok, lo_k, mid)
    return i2f_np(hi_k)


# ---------------------------------------------------------------- nf8
levels8 = quantize_8bit(torch.ones(2), quant_type="nf8")[1].clone()


def nf8_decide(v):
    # the reference's element decision on an already normalized value: abs_max = 1 keeps v / abs_max == v
    x = torch.cat([v.reshape(-1), torch.tensor([1.0])])
    idx = quantize_8bit(x, quant_type="nf8")[0]
    return idx[:-1].to(torch.int64).reshape(v.shape)


t8 = bisect(nf8_decide, levels8[:-1].clone(), levels8[1:].clone(), torch.arange(1, 256))
# monotone around every threshold
for d in range(-16, 17):
    probe = i2f_np(f2i(t8) + d)
    want = torch.arange(1, 256) - (1 if d < 0 else 0)
    assert bool((nf8_decide(probe) == want).all()), d

# ---------------------------------------------------------------- fp4 / fp8 exponent fields
def exp_field(fn, qt, shift, mask):
    def decide(a):
        x = a.reshape(-1)
        q = fn(x, quant_type=qt)[0]
        return ((q.to(torch.int64) >> shift) & mask).reshape(a.shape)
    return decide


def exp_thresholds(fn, qt, shift, mask, bias, emax):
    decide = exp_field(fn, qt, shift, mask)
    k = torch.arange(1, emax + 1)
    lo = torch.pow(torch.tensor(2.0), (k - bias - 1).float())          # 2^(k-1-bias): field k-1
    hi = torch.pow(torch.tensor(2.0), (k - bias).float())              # 2^(k-bias): field k
    t = bisect(decide, lo, hi, k)
    for d in range(-16, 17):
        probe = i2f_np(f2i(t) + d)
        want = k - (1 if d < 0 else 0)
        assert bool((decide(probe) == want).all()), (qt, d)
    return t


t_fp4 = exp_thresholds(quantize_4bit, "fp4", 1, 0x3, 1, 3)
t_fp8 = exp_thresholds(quantize_8bit, "fp8", 3, 0xF, 7, 15)
# 2 ** (e - bias) must be the exact power of two for every field value
for bias, emax in ((1, 3), (7, 15)):
    e = torch.arange(0, emax + 1).float()
    assert bool((2 ** (e - bias) == torch.ldexp(torch.ones(emax + 1), (e - bias).to(torch.int32))).all())

np.savez(os.path.join(HERE, "quanta_tables_n4.npz"), nf8_levels=levels8.numpy(), nf8_thresholds=t8.numpy(),
         fp4_exp_thresholds=t_fp4.numpy(), fp8_exp_thresholds=t_fp8.numpy())
print("nf8 thresholds", t8[:3].tolist(), "...", "fp4", t_fp4.tolist(), "fp8", [hex(v) for v in t_fp8.view(torch.int32).tolist()])

# ---------------------------------------------------------------- golden vectors
g = torch.Generator().manual_seed(2468)
store, manifest = {}, []


def add(qt, x):
    name = f"{qt}_{len(manifest):03d}"
    if qt == "fp4":
        q, _, bias = quantize_4bit(x, quant_type="fp4")
        deq = dequantize_4bit(q, None, bias, quant_type="fp4")
    elif qt == "fp8":
        q, _, bias = quantize_8bit(x, quant_type="fp8")
        deq = dequantize_8bit(q, None, bias, quant_type="fp8")
    else:
        q, lv, am = quantize_8bit(x, quant_type="nf8")
        deq = dequantize_8bit(q, lv, am, quant_type="nf8")
        store[f"{name}/absmax"] = am.numpy().reshape(())
    store[f"{name}/x"] = x.numpy()
    store[f"{name}/q"] = q.numpy()
    store[f"{name}/deq"] = deq.numpy()
    manifest.append({"name": name, "kind": qt})


base = [torch.tensor([-1.0, 0.0, 1.0, 2.0]), torch.randn(4096, generator=g), torch.randn(33, 65, generator=g) * 0.02,
        torch.randn(2048, generator=g) * 37.0, torch.zeros(64), torch.full((64,), -2.0),
        torch.randn(512, generator=g) * 1e-20, torch.randn(512, generator=g) * 1e20,
        torch.tensor([0.0, -0.0, 1e-45, -1e-45, 1.1754944e-38, 3.4028235e38, -3.4028235e38, 0.5, 0.75, 1.5, 3.0, 6.0, 12.0])]
for qt, t in (("fp4", t_fp4), ("fp8", t_fp8)):
    edge = torch.cat([i2f_np(f2i(t) + d) for d in range(-3, 4)])
    # mantissa rounding boundaries: (1 + (m + 0.5) / M) * 2^k and neighbours
    M = 1 if qt == "fp4" else 8
    bias, emax = (1, 3) if qt == "fp4" else (7, 15)
    ks = torch.arange(-bias - 2, emax - bias + 3).float()
    mant = torch.cat([(1 + (m + 0.5) / M) * 2 ** ks for m in range(-1, M + 1)])
    mant = torch.cat([i2f_np(f2i(mant) + d) for d in (-1, 0, 1)])
    for x in base + [torch.cat([edge, -edge]), torch.cat([mant, -mant])]:
        add(qt, x)
edge8 = torch.cat([i2f_np(f2i(t8) + d) for d in range(-2, 3)] + [levels8])
for x in base[:8] + [edge8 * 1.0, torch.cat([edge8 * 2.5, torch.tensor([2.5])]), torch.cat([edge8 * 0.37, torch.tensor([-0.37])])]:
    add("nf8", x)
t = torch.randn(192, generator=g); t[5] = float("inf")
add("nf8", t)
store["manifest"] = np.array(json.dumps(manifest))
np.savez_compressed(os.path.join(HERE, "quanta_golden_n4.npz"), **store)
print(len(manifest), "golden cases")
