"""NF4 (row N1): the numpy oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden_nf4.py)."""
import json
import os

import numpy as np

from oracle import oracle_np as O

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quanta_golden_nf4.npz"))
MANIFEST = json.loads(bytes(Z["manifest"]).decode())


def same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def test_levels_are_the_reference_table():
    assert np.array_equal(O.NF4_LEVELS.view(np.uint32), Z["levels"].view(np.uint32))


def test_nf4_oracle_matches_reference_golden():
    assert len(MANIFEST) >= 20
    for c in MANIFEST:
        x = Z[f"{c['name']}/x"]
        idx, am = O.quantize_nf4(x, c["block"])
        assert np.array_equal(idx, Z[f"{c['name']}/idx"]), c
        assert same_bits(am, Z[f"{c['name']}/absmax"]), c
        assert same_bits(O.dequantize_nf4(idx, am, c["block"]), Z[f"{c['name']}/deq"]), c
