"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called
through the reference-shaped Python API (which goes through the C-ABI), must
reproduce the reference's golden outputs and the oracle bit for bit."""
import numpy as np
import pytest
import torch

from helpers import assert_f32_bits, assert_u8_equal, ulp_diff
from oracle import oracle_c as OC
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu

MODES = {"tensor": 0, "dim0": 1, "block": 2}


@pytest.fixture(scope="module")
def Q():
    import quanta_b200
    assert torch.cuda.is_available()
    return quanta_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def quantize(Q, x, bits, mode, block=0, **kw):
    fn = Q.quantize_8bit if bits == 8 else Q.quantize_4bit
    if mode == 0:
        return fn(x, **kw)
    if mode == 1:
        return fn(x, per_channel=True, **kw)
    return fn(x, blocksize=block, **kw)


# ---------------------------------------------------------------- golden vectors

@pytest.mark.parametrize("row", ["A_tensor", "A_dim0", "A_block"])
def test_convention_a_golden(Q, golden, row):
    for c in golden.cases(row):
        kw = c["kwargs"]
        x = golden.x(c)
        q, s, z = quantize(Q, dev(x), kw["bits"], MODES[kw["mode"]], kw.get("block", 0))
        what = f"{c['name']} {kw} shape={x.shape}"
        assert q.dtype == torch.uint8 and tuple(q.shape) == x.shape and q.is_cuda
        assert s.dtype == torch.float32 and tuple(s.shape) == golden.get(c, "scale").shape, what
        assert_u8_equal(host(q), golden.get(c, "q"), what + " codes")
        assert_f32_bits(host(s), golden.get(c, "scale"), what + " scale")
        assert_f32_bits(host(z), golden.get(c, "zp"), what + " zp", zero_sign_free=True)
        dq = Q.dequantize_8bit if kw["bits"] == 8 else Q.dequantize_4bit
        d = dq(dev(golden.get(c, "q")), dev(golden.get(c, "scale")), dev(golden.get(c, "zp")),
               **({"blocksize": kw["block"]} if kw["mode"] == "block" else {}))
        assert_f32_bits(host(d), golden.get(c, "deq"), what + " dequant")


def test_convention_b_golden(Q, golden):
    from quanta_b200.backends import quantize_8bit, quantize_4bit, dequantize_8bit, dequantize_4bit
    for c in golden.cases("B"):
        kw = c["kwargs"]
        x = golden.x(c)
        qf, df = (quantize_8bit, dequantize_8bit) if kw["bits"] == 8 else (quantize_4bit, dequantize_4bit)
        q, s, z = qf(dev(x), kw["per_channel"], kw["symmetric"])
        what = f"{c['name']} {kw} shape={x.shape}"
        assert tuple(s.shape) == golden.get(c, "scale").shape, what
        assert_u8_equal(host(q), golden.get(c, "q"), what + " codes")
        assert_f32_bits(host(s), golden.get(c, "scale"), what + " scale")
        assert_f32_bits(host(z), golden.get(c, "zp"), what + " zp", zero_sign_free=True)
        d = df(dev(golden.get(c, "q")), dev(golden.get(c, "scale")), dev(golden.get(c, "zp")))
        assert_f32_bits(host(d), golden.get(c, "deq"), what + " dequant")


def test_pack_unpack_golden(Q, golden):
    for c in golden.cases("P"):
        qv = golden.get(c, "q")
        packed, shape = Q.pack_4bit_tensor(dev(qv))
        assert tuple(shape) == qv.shape
        assert_u8_equal(host(packed), golden.get(c, "packed"), c["name"] + " pack")
        assert_u8_equal(host(Q.unpack_4bit_tensor(dev(golden.get(c, "packed")))), golden.get(c, "unpacked"),
                        c["name"] + " unpack")


# ---------------------------------------------------------------- oracle parity, random inputs

SHAPES = [(1, 32), (3, 64), (257, 96), (1024, 1024), (300, 1000), (77, 4096)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("bits", [8, 4])
def test_tensor_and_dim0_vs_oracle(Q, shape, bits):
    rng = np.random.default_rng(hash((shape, bits)) % 2**32)
    x = (rng.standard_normal(shape) * 0.02).astype(np.float32)
    for mode in (0, 1):
        q, s, z = quantize(Q, dev(x), bits, mode)
        qo, so, zo = OC.quantize_affine(x, bits, mode)
        assert_u8_equal(host(q), qo, f"codes mode={mode}")
        assert_f32_bits(host(s), so, "scale")
        assert_f32_bits(host(z), zo, "zp")
        d = (Q.dequantize_8bit if bits == 8 else Q.dequantize_4bit)(q, s, z)
        assert_f32_bits(host(d), OC.dequantize_affine(qo, so, zo, mode, x.shape[1]), "dequant")


@pytest.mark.parametrize("shape", [(2000, 4096), (4100, 2048), (8192, 1024), (130, 8192), (513, 32)])
@pytest.mark.parametrize("bits", [8, 4])
def test_large_tensor_and_dim0_single_launch(Q, shape, bits):
    """Sizes at which the single-launch kernels keep some tiles in shared memory / registers and
    re-stream the rest (more tiles per CTA than slots), ragged last row tile, constant columns."""
    rng = np.random.default_rng(shape[0] * 7 + bits)
    x = (rng.standard_normal(shape) * 0.02).astype(np.float32)
    x[:, 5] = 0.25                                      # constant column -> +1e-6 rule
    x[:, 17] = 0.0
    x[3, 30] = 40.0                                     # wide-range column
    for mode in (0, 1):
        q, s, z = quantize(Q, dev(x), bits, mode)
        qo, so, zo = OC.quantize_affine(x, bits, mode)
        assert_u8_equal(host(q), qo, f"codes mode={mode}")
        assert_f32_bits(host(s), so, "scale")
        assert_f32_bits(host(z), zo, "zp")
    if bits == 4:
        pk, s2, z2 = Q.quantize_4bit(dev(x), per_channel=True, packed=True)
        qo, so, zo = OC.quantize_affine(x, 4, 1)
        assert_u8_equal(host(pk), O.pack4(qo.reshape(-1)), "dim0 fused pack == pack(unpacked)")
    xb = torch.from_numpy(x).to(torch.bfloat16)
    for mode in (0, 1):
        q, s, z = quantize(Q, xb.cuda(), bits, mode)
        qo, so, zo = OC.quantize_affine(xb.float().numpy(), bits, mode)
        assert_u8_equal(host(q), qo, f"bf16 codes mode={mode}")
        assert_f32_bits(host(s), so, "bf16 scale")


@pytest.mark.parametrize("block", [32, 64, 128, 256, 512, 1024, 2048, 16, 96, 6])
@pytest.mark.parametrize("bits", [8, 4])
def test_blockwise_vs_oracle(Q, block, bits):
    rng = np.random.default_rng(block * 10 + bits)
    n = block * 1531                                   # odd number of blocks: partial TMA tile
    x = (rng.standard_normal(n) * 0.02).astype(np.float32)
    x[:block] = 0.5                                    # constant block -> +1e-6 rule
    x[block:2 * block] = 0.0
    q, s, z = quantize(Q, dev(x), bits, 2, block)
    qo, so, zo = OC.quantize_affine(x, bits, 2, block)
    assert_u8_equal(host(q), qo, "codes")
    assert_f32_bits(host(s), so, "scale")
    assert_f32_bits(host(z), zo, "zp")
    d = (Q.dequantize_8bit if bits == 8 else Q.dequantize_4bit)(q, s, z, blocksize=block)
    assert_f32_bits(host(d), OC.dequantize_affine(qo, so, zo, 2, block), "dequant")
    if bits == 4:
        pk, s2, z2 = Q.quantize_4bit(dev(x), blocksize=block, packed=True)
        assert_u8_equal(host(pk), O.pack4(qo), "fused pack == pack(unpacked)")
        assert_f32_bits(host(s2), so, "scale (packed)")
        d = Q.dequantize_4bit(pk, s2, z2, blocksize=block, packed=True, shape=x.shape)
        assert_f32_bits(host(d), OC.dequantize_affine(qo, so, zo, 2, block), "packed dequant")
        for od, name in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
            d16 = Q.dequantize_4bit(pk, s2, z2, blocksize=block, packed=True, shape=x.shape, out_dtype=od)
            want = torch.from_numpy(OC.dequantize_affine(qo, so, zo, 2, block)).to(od)
            assert torch.equal(d16.cpu(), want), name


def test_config2_shapes_vs_oracle(Q):
    """One Llama-2-7B layer's matrices (config 2), 4-bit block-64 quantize+pack."""
    for i, shape in enumerate([(4096, 4096), (11008, 4096), (4096, 11008)]):
        g = torch.Generator().manual_seed(1234 + i)
        x = (torch.randn(shape, generator=g) * 0.02)
        pk, s, z = Q.quantize_4bit(x.cuda(), blocksize=64, packed=True)
        po, so, zo = OC.quantize4_block_pack(x.numpy(), 64)
        assert_u8_equal(host(pk), po, f"packed {shape}")
        assert_f32_bits(host(s), so, "scale")
        assert_f32_bits(host(z), zo, "zp")


def test_config1_roundtrip_vs_oracle(Q):
    """Config 1: 4096x4096 fp32 quantize_8bit / dequantize_8bit round trip."""
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4096, 4096, generator=g)
    q, s, z = Q.quantize_8bit(x.cuda())
    qo, so, zo = OC.quantize_affine(x.numpy(), 8, 0)
    assert s.dim() == 0 and z.dim() == 0
    assert_u8_equal(host(q), qo, "codes")
    assert_f32_bits(host(s), so, "scale")
    assert_f32_bits(host(z), zo, "zp")
    d = Q.dequantize_8bit(q, s, z)
    assert_f32_bits(host(d), OC.dequantize_affine(qo, so, zo, 0), "dequant")
    assert torch.allclose(d.cpu(), x, atol=float(s) * 0.5001)


# ---------------------------------------------------------------- edge cases

def test_near_tie_division_is_exact(Q):
    """Adversarial inputs: (x - min)/scale within a few ulps of k + 0.5 for many
    scales — the hoisted-reciprocal divide must match a true IEEE divide."""
    rng = np.random.default_rng(5)
    nb = 20000
    for bits, L in ((4, 15), (8, 255)):
        x = np.zeros((nb, 64), np.float32)
        s = np.ldexp(1 + rng.random(nb), rng.integers(-40, 40, nb)).astype(np.float32)
        x[:, 1] = (s * np.float32(L)).astype(np.float32)
        scale = ((x[:, 1] - x[:, 0]) / np.float32(L)).astype(np.float32)
        k = rng.integers(0, L, (nb, 62)).astype(np.float32)
        t = ((k + np.float32(0.5)) * scale[:, None]).astype(np.float32)
        t = (t.view(np.int32) + rng.integers(-3, 4, t.shape).astype(np.int32)).view(np.float32)
        x[:, 2:] = np.minimum(t, x[:, 1:2])
        q, sc, z = quantize(Q, dev(x.reshape(-1)), bits, 2, 64)
        qo, so, zo = OC.quantize_affine(x.reshape(-1), bits, 2, 64)
        assert_f32_bits(host(sc), so, "scale")
        assert_u8_equal(host(q), qo, f"near-tie codes bits={bits}")


def test_nonfinite_and_extreme_inputs(Q):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((64, 64)).astype(np.float32)
    x[0, 3] = np.nan
    x[1, 5] = np.inf
    x[2, 7] = -np.inf
    x[3] = 1000.0                         # |mn| >= 32: mn + 1e-6 == mn -> scale 0 -> NaN -> code 0
    x[4] = 1e-30 * rng.standard_normal(64)
    x[5] = 1e30 * rng.standard_normal(64)
    x[6] = 1e-42                          # denormals
    x[7, :32] = -0.0
    for bits in (8, 4):
        q, s, z = quantize(Q, dev(x.reshape(-1)), bits, 2, 64)
        qo, so, zo = OC.quantize_affine(x.reshape(-1), bits, 2, 64)
        assert_u8_equal(host(q), qo, "codes")
        assert_f32_bits(host(s), so, "scale")
        assert_f32_bits(host(z), zo, "zp")
        for mode in (0, 1):
            q, s, z = quantize(Q, dev(x), bits, mode)
            qo, so, zo = OC.quantize_affine(x, bits, mode)
            assert_u8_equal(host(q), qo, f"codes mode={mode}")
            assert_f32_bits(host(s), so, "scale")
    y = x.copy()
    y[np.isnan(y)] = 0
    for mode in (0, 1):
        for bits in (8, 4):
            q, s, z = quantize(Q, dev(y), bits, mode)
            qo, so, zo = OC.quantize_affine(y, bits, mode)
            assert_u8_equal(host(q), qo, f"codes mode={mode} (inf)")


def test_reference_edge_cases(Q):
    """Quanta/tests/test_quantization.py:114-124 (CUDA twins of the edge cases)."""
    q, s, z = Q.quantize_8bit(torch.zeros(5).cuda())
    assert torch.all(q == 0)
    q, s, z = Q.quantize_8bit((torch.ones(5) * 2.0).cuda())
    assert torch.all(q == q[0])
    assert host(s).view(np.uint32) == 0x31808081 and float(z) == 2.0


def test_strided_unaligned_and_odd_inputs(Q):
    rng = np.random.default_rng(3)
    base = torch.from_numpy(rng.standard_normal(5000).astype(np.float32)).cuda()
    for off, n in ((1, 4097), (3, 333), (0, 31), (2, 1)):
        x = base[off:off + n]                     # 4-byte aligned only -> generic path
        xn = host(x)
        for bits in (8, 4):
            q, s, z = quantize(Q, x, bits, 0)
            qo, so, zo = OC.quantize_affine(xn, bits, 0)
            assert_u8_equal(host(q), qo, f"codes off={off} n={n}")
            assert_f32_bits(host(s), so, "scale")
        if n % 2 == 1:
            pk, _, _ = Q.quantize_4bit(x, packed=True)
            assert_u8_equal(host(pk), O.pack4(OC.quantize_affine(xn, 4, 0)[0]), "odd-length fused pack")
    m = torch.from_numpy(rng.standard_normal((64, 130)).astype(np.float32)).cuda()
    xt = m.t()                                     # non-contiguous view, like the reference accepts
    for bits in (8, 4):
        q, s, z = quantize(Q, xt, bits, 1)
        qo, so, zo = OC.quantize_affine(host(xt), bits, 1)
        assert_u8_equal(host(q), qo, "codes (transposed view, odd cols)")
        assert_f32_bits(host(s), so, "scale")
    x4 = torch.from_numpy(rng.standard_normal((6, 3, 5, 4)).astype(np.float32)).cuda()
    q, s, z = Q.quantize_8bit(x4, per_channel=True)
    assert tuple(s.shape) == (1, 3, 5, 4)
    qo, so, zo = OC.quantize_affine(host(x4), 8, 1)
    assert_u8_equal(host(q), qo, "4-D per_channel")
    assert_f32_bits(host(Q.dequantize_8bit(q, s, z)), OC.dequantize_affine(qo, so, zo, 1, 60), "4-D dequant")


def test_half_inputs_follow_fp32_arithmetic(Q):
    """fp16 / bf16 inputs are widened to fp32 (declared deviation)."""
    rng = np.random.default_rng(8)
    x32 = torch.from_numpy((rng.standard_normal((256, 512)) * 0.02).astype(np.float32))
    for dt in (torch.float16, torch.bfloat16):
        x = x32.to(dt)
        xo = x.float().numpy()
        for bits in (8, 4):
            q, s, z = quantize(Q, x.cuda(), bits, 2, 64)
            qo, so, zo = OC.quantize_affine(xo.reshape(-1), bits, 2, 64)
            assert_u8_equal(host(q).reshape(-1), qo, f"block codes {dt}")
            assert_f32_bits(host(s), so, "scale")
            for mode in (0, 1):
                q, s, z = quantize(Q, x.cuda(), bits, mode)
                qo, so, zo = OC.quantize_affine(xo, bits, mode)
                assert_u8_equal(host(q), qo, f"codes {dt} mode={mode}")
                assert_f32_bits(host(s), so, "scale")


def test_backend_vs_oracle_random(Q):
    from quanta_b200.backends import quantize_8bit, quantize_4bit, dequantize_8bit, dequantize_4bit
    rng = np.random.default_rng(21)
    for shape, mul, add in (((512, 384), 1.0, 0.0), ((1000, 36), 0.02, 0.3), ((64, 4096), 5.0, -2.0), ((4097,), 1.0, 0.0)):
        x = (rng.standard_normal(shape) * mul + add).astype(np.float32)
        for bits, qf, df in ((8, quantize_8bit, dequantize_8bit), (4, quantize_4bit, dequantize_4bit)):
            for pc in (False, True):
                if pc and x.ndim < 2:
                    continue
                for sym in (True, False):
                    q, s, z = qf(dev(x), pc, sym)
                    qo, so, zo = OC.backend_quantize(x, bits, pc, sym)
                    what = f"B shape={shape} bits={bits} pc={pc} sym={sym}"
                    assert_u8_equal(host(q), qo, what)
                    assert_f32_bits(host(s), so, what + " scale")
                    assert_f32_bits(host(z), zo, what + " zp")
                    assert_f32_bits(host(df(q, s, z)), OC.backend_dequantize(qo, so, zo, bits), what + " dequant")


def test_pack_unpack_random(Q):
    rng = np.random.default_rng(2)
    for n in (1, 2, 15, 16, 17, 4096, 100003, 1 << 22):
        qv = rng.integers(0, 256, n).astype(np.uint8)          # includes unmasked values > 15
        packed, shape = Q.pack_4bit_tensor(dev(qv))
        assert_u8_equal(host(packed), OC.pack4(qv), f"pack n={n}")
        assert_u8_equal(host(Q.unpack_4bit_tensor(packed)), OC.unpack4(OC.pack4(qv)), f"unpack n={n}")
    base = dev(rng.integers(0, 16, 1001).astype(np.uint8))
    v = base[1:]                                                # unaligned view
    assert_u8_equal(host(Q.pack_4bit_tensor(v)[0]), OC.pack4(host(v)), "unaligned pack")


# ---------------------------------------------------------------- full-size properties

def test_full_size_properties(Q):
    """Largest config-2 matrix, checked through size-independent properties:
    idempotence, unpack(packed) == unpacked codes, round trip within scale/2."""
    torch.manual_seed(0)
    x = torch.randn(11008, 4096, device="cuda") * 0.02
    q, s, z = Q.quantize_4bit(x, blocksize=64)
    pk, s2, z2 = Q.quantize_4bit(x, blocksize=64, packed=True)
    assert torch.equal(s, s2) and torch.equal(z, z2)
    assert torch.equal(Q.unpack_4bit_tensor(pk), q.reshape(-1))
    assert torch.equal(Q.pack_4bit_tensor(q)[0], pk)
    assert int(q.max()) == 15 and int(q.min()) == 0
    d = Q.dequantize_4bit(pk, s, z, blocksize=64, packed=True, shape=x.shape)
    err = (d - x).abs().reshape(-1, 64)
    assert bool((err <= s[:, None] * 0.5001 + 1e-12).all())
    # every block attains code 0 at its min and code 15 at its max
    qb = q.reshape(-1, 64)
    assert bool((qb.min(dim=1).values == 0).all()) and bool((qb.max(dim=1).values == 15).all())
    # re-quantizing the dequantized tensor reproduces the codes (grid points are fixed points)
    q3, _, _ = Q.quantize_4bit(d, blocksize=64)
    assert (q3 != q).float().mean().item() < 1e-3


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_batched_blockwise_dequantize_equals_per_tensor_calls(out_dtype):
    """quanta_dequantize_block_batch: many tensors per launch, bit-identical to one dequantize call per tensor (and, in
    fp32, to the oracle); > 16 tensors exercises several launches, sizes below one chunk and ragged last chunks."""
    import quanta_b200 as Q
    g = torch.Generator().manual_seed(22)
    shapes = [(256, 512), (64, 64), (1000, 64), (4096, 1024), (32, 8192), (3, 64)] * 3 + [(11008, 512)]
    ts = [(torch.randn(*s, generator=g) * 0.02).cuda() for s in shapes]
    for bits, packed in ((4, True), (4, False), (8, False)):
        quant = Q.quantize_4bit_many(ts, blocksize=64, packed=packed) if bits == 4 else Q.quantize_8bit_many(ts, blocksize=64)
        qs, ss, zs = [t[0] for t in quant], [t[1] for t in quant], [t[2] for t in quant]
        if bits == 4:
            many = Q.dequantize_4bit_many(qs, ss, zs, blocksize=64, packed=packed, shapes=[t.shape for t in ts] if packed else None,
                                          out_dtype=out_dtype)
        else:
            many = Q.dequantize_8bit_many(qs, ss, zs, blocksize=64, out_dtype=out_dtype)
        assert len(many) == len(ts)
        for t, q, sc, z, d in zip(ts, qs, ss, zs, many):
            one = (Q.dequantize_4bit(q, sc, z, blocksize=64, packed=packed, shape=t.shape if packed else None, out_dtype=out_dtype)
                   if bits == 4 else Q.dequantize_8bit(q, sc, z, blocksize=64, out_dtype=out_dtype))
            assert d.shape == t.shape and d.dtype == out_dtype
            assert torch.equal(d.view(torch.int16 if out_dtype != torch.float32 else torch.int32),
                               one.view(torch.int16 if out_dtype != torch.float32 else torch.int32))
    if out_dtype == torch.float32:
        q, sc, z = Q.quantize_8bit(ts[3], blocksize=64)
        d = Q.dequantize_8bit_many([q], [sc], [z], blocksize=64)[0]
        ref = O.dequantize_affine(q.cpu().numpy().reshape(-1, 64), sc.cpu().numpy().reshape(-1, 1), z.cpu().numpy().reshape(-1, 1))
        assert np.array_equal(d.cpu().numpy().reshape(-1, 64).view(np.uint32), np.asarray(ref, dtype=np.float32).view(np.uint32))
    with pytest.raises(ValueError):
        Q.dequantize_8bit_many([qs[0]], [ss[0]], [zs[0], zs[1]], blocksize=64)


def test_batched_blockwise_quantize_equals_per_tensor_calls():
    """quanta_quantize_block_batch: many tensors per launch, bit-identical to one call per tensor
    (and therefore to the oracle); > 16 tensors exercises several launches, odd sizes the fallback."""
    import quanta_b200 as Q
    g = torch.Generator().manual_seed(21)
    shapes = [(256, 512), (64, 64), (1000, 64), (4096, 1024), (32, 8192)] * 4 + [(3, 64)]
    ts = [(torch.randn(*s, generator=g) * 0.02).cuda() for s in shapes]
    for bits, packed in ((4, True), (4, False), (8, False)):
        many = Q.quantize_4bit_many(ts, blocksize=64, packed=packed) if bits == 4 else Q.quantize_8bit_many(ts, blocksize=64)
        assert len(many) == len(ts)
        for t, (q, s, z) in zip(ts, many):
            q1, s1, z1 = (Q.quantize_4bit(t, blocksize=64, packed=packed) if bits == 4 else Q.quantize_8bit(t, blocksize=64))
            assert torch.equal(q, q1) and torch.equal(s.view(torch.int32), s1.view(torch.int32))
            assert torch.equal(z.view(torch.int32), z1.view(torch.int32))
    # against the oracle directly for one tensor
    po, so, zo = O.quantize4_block_pack(ts[3].cpu().numpy(), 64)
    q, s, z = Q.quantize_4bit_many(ts, blocksize=64, packed=True)[3]
    assert np.array_equal(q.cpu().numpy(), po) and np.array_equal(s.cpu().numpy().view(np.uint32), so.view(np.uint32))


@pytest.mark.parametrize("mode", ["tensor", "dim0", "block", "many"])
def test_repeated_calls_are_identical(Q, mode):
    """Race regression: a TMA stage used to be released before the shared-memory loads of its rows had
    returned, so the next tile occasionally overwrote 16-byte chunks that were still to be read (~3 % of
    per-tensor calls at this size).  Every repetition must reproduce the first result bit for bit."""
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(6144, 4096, device="cuda", generator=g)
    if mode == "tensor":
        fn = lambda: Q.quantize_8bit(x)[0]
    elif mode == "dim0":
        fn = lambda: Q.quantize_8bit(x, per_channel=True)[0]
    elif mode == "block":
        fn = lambda: Q.quantize_4bit(x, blocksize=64, packed=True)[0]
    else:
        parts = [x[:2048], x[2048:4096], x[4096:]]
        fn = lambda: torch.cat([t[0] for t in Q.quantize_4bit_many(parts, blocksize=64, packed=True)])
    ref = fn().clone()
    for _ in range(60):
        assert torch.equal(fn(), ref)
