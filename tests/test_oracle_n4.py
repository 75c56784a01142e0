"""nf8 / fp4 / fp8 (row N4): the numpy oracle against golden vectors produced by the unmodified reference
(tests/golden/make_tables_n4.py)."""
import json
import os

import numpy as np

from oracle import oracle_np as O

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quanta_golden_n4.npz"))
MANIFEST = json.loads(str(Z["manifest"]))


def same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def test_golden_covers_all_three_formats():
    kinds = {c["kind"] for c in MANIFEST}
    assert kinds == {"fp4", "fp8", "nf8"} and len(MANIFEST) >= 30


def test_fp4_fp8_oracle_matches_reference_golden():
    for c in MANIFEST:
        if c["kind"] not in ("fp4", "fp8"):
            continue
        bits, bias = (4, 1) if c["kind"] == "fp4" else (8, 7)
        x = Z[f"{c['name']}/x"]
        q = O.quantize_fp(x, bits)
        assert np.array_equal(q, Z[f"{c['name']}/q"]), c
        assert same_bits(O.dequantize_fp(q, bits, bias), Z[f"{c['name']}/deq"]), c


def test_nf8_oracle_matches_reference_golden():
    for c in MANIFEST:
        if c["kind"] != "nf8":
            continue
        x = Z[f"{c['name']}/x"]
        idx, am = O.quantize_nf8(x)
        assert np.array_equal(idx, Z[f"{c['name']}/q"]), c
        assert same_bits(am, Z[f"{c['name']}/absmax"]), c
        assert same_bits(O.dequantize_nf8(idx, am), Z[f"{c['name']}/deq"]), c


def test_known_answer_vectors():
    """[-1, 0, 1, 2] -> fp4 [10, 2, 2, 4] -> [-1, 1, 1, 2]; fp8 [184, 56, 56, 64]."""
    x = np.array([-1.0, 0.0, 1.0, 2.0], np.float32)
    assert O.quantize_fp(x, 4).tolist() == [10, 2, 2, 4]
    assert O.dequantize_fp(O.quantize_fp(x, 4), 4, 1).tolist() == [-1.0, 1.0, 1.0, 2.0]
    assert O.quantize_fp(x, 8).tolist() == [184, 56, 56, 64]
    assert O.dequantize_fp(O.quantize_fp(x, 8), 8, 7).tolist() == [-1.0, 1.0, 1.0, 2.0]
