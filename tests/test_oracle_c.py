"""The C oracle (oracle/quanta_oracle.c) against the reference's golden vectors
and against the numpy oracle on random inputs.  CPU only."""
import numpy as np
import pytest

from oracle import oracle_np as O
from oracle import oracle_c as OC
from helpers import assert_f32_bits, assert_u8_equal

MODES = {"tensor": 0, "dim0": 1, "block": 2}


@pytest.mark.parametrize("row", ["A_tensor", "A_dim0", "A_block"])
def test_c_convention_a_golden(golden, row):
    for c in golden.cases(row):
        kw = c["kwargs"]
        x = golden.x(c)
        mode = MODES[kw["mode"]]
        q, s, z = OC.quantize_affine(x, kw["bits"], mode, kw.get("block", 0))
        what = f"{c['name']} {kw}"
        assert_u8_equal(q, golden.get(c, "q"), what)
        assert_f32_bits(s, golden.get(c, "scale"), what + " scale")
        assert_f32_bits(z, golden.get(c, "zp"), what + " zp", zero_sign_free=True)
        p = {0: 0, 1: x.size // x.shape[0], 2: kw.get("block", 0)}[mode]
        d = OC.dequantize_affine(golden.get(c, "q"), golden.get(c, "scale"), golden.get(c, "zp"), mode, p)
        assert_f32_bits(d, golden.get(c, "deq"), what + " dequant")
        if kw["bits"] == 4 and mode == 2 and x.size % 2 == 0:
            pk, s2, z2 = OC.quantize4_block_pack(x, kw["block"])
            assert_u8_equal(pk, O.pack4(golden.get(c, "q")), what + " fused pack")
            assert_f32_bits(s2, golden.get(c, "scale"), what + " fused scale")
            d = OC.dequantize_affine(pk, s2, z2, mode, p, packed=True)
            assert_f32_bits(d, golden.get(c, "deq"), what + " packed dequant")


def test_c_convention_b_golden(golden):
    for c in golden.cases("B"):
        kw = c["kwargs"]
        x = golden.x(c)
        q, s, z = OC.backend_quantize(x, kw["bits"], kw["per_channel"], kw["symmetric"])
        what = f"{c['name']} {kw}"
        assert_u8_equal(q, golden.get(c, "q"), what)
        assert_f32_bits(s, golden.get(c, "scale"), what + " scale")
        assert_f32_bits(z, golden.get(c, "zp"), what + " zp", zero_sign_free=True)
        d = OC.backend_dequantize(golden.get(c, "q"), golden.get(c, "scale"), golden.get(c, "zp"), kw["bits"])
        assert_f32_bits(d, golden.get(c, "deq"), what + " dequant")


def test_c_pack_golden(golden):
    for c in golden.cases("P"):
        assert_u8_equal(OC.pack4(golden.get(c, "q")), golden.get(c, "packed"), c["name"])
        assert_u8_equal(OC.unpack4(golden.get(c, "packed")), golden.get(c, "unpacked"), c["name"])


@pytest.mark.parametrize("seed", range(4))
def test_c_vs_numpy_random(seed):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((96, 256)) * [0.02, 1.0, 30.0, 1e-4][seed] + [0, 0, 5, 0][seed]).astype(np.float32)
    x[3, :64] = 1.5                      # constant block
    x[5, 7] = np.nan if seed == 1 else x[5, 7]
    x[9, 9] = np.inf if seed == 2 else x[9, 9]
    for bits in (8, 4):
        for mode, blk in ((0, 0), (1, 0), (2, 64), (2, 128)):
            a = OC.quantize_affine(x, bits, mode, blk)
            b = O.quantize_affine(x, bits, mode, blk)
            assert_u8_equal(a[0], b[0], f"codes bits={bits} mode={mode}")
            assert_f32_bits(a[1], b[1], "scale")
            assert_f32_bits(a[2], b[2], "zp")
        for pc in (False, True):
            for sym in (True, False):
                a = OC.backend_quantize(x, bits, pc, sym)
                b = O.backend_quantize(x, bits, pc, sym)
                assert_u8_equal(a[0], b[0], f"B codes bits={bits} pc={pc} sym={sym}")
                assert_f32_bits(a[1], b[1], "B scale")
                assert_f32_bits(a[2], b[2], "B zp")
                assert_f32_bits(OC.backend_dequantize(*a, bits=bits), O.backend_dequantize(*b, bits=bits), "B deq")


def test_c_linear_dequant_vs_numpy():
    rng = np.random.default_rng(7)
    N, K, M = 48, 256, 5
    w = (rng.standard_normal((N, K)) * 0.02).astype(np.float32)
    x = O._round_to(rng.standard_normal((M, K)).astype(np.float32), "bf16")
    bias = rng.standard_normal(N).astype(np.float32)
    for bits in (8, 4):
        q, s, z = O.quantize_affine(w, bits, O.MODE_BLOCK, 64)
        ref = O.linear_dequant(x, q, s, z, bias, 64, "bf16")
        wq = q if bits == 8 else O.pack4(q)
        y = OC.linear_dequant(x, wq, s, z, bias, bits, N, K, 64)
        np.testing.assert_allclose(y, ref, rtol=1e-6, atol=1e-6)
