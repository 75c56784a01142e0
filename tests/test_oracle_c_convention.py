"""Convention C (BaseQuantizer, row N4): the numpy oracle against the golden vectors produced by the unmodified
reference (tests/golden/make_golden_c.py), bit for bit."""
import json
import os

import numpy as np

from oracle import oracle_np as O

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quanta_golden_c.npz"))
CASES = json.loads(bytes(Z["manifest"]).decode())


def same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_oracle_matches_reference_golden():
    assert len(CASES) >= 90
    for c in CASES:
        x = Z[c["name"] + "/x"]
        q, s, z = O.base_quantize(x, c["bits"], c["per_channel"], c["symmetric"])
        assert np.array_equal(q, Z[c["name"] + "/q"]), c
        assert same_bits(s, Z[c["name"] + "/scale"]) and same_bits(z, Z[c["name"] + "/zp"]), c
        d = O.base_dequantize(q, s, z, c["bits"], c["symmetric"])
        assert same_bits(d, Z[c["name"] + "/deq"]), c


def test_appendix_b_known_answers():
    q, s, _ = O.base_quantize(np.array([-1, -.5, 0, .5, 1], np.float32), 8, False, True)
    assert q.tolist() == [1, 64, 128, 192, 255] and float(s) == 127.0
    q, s, _ = O.base_quantize(np.array([-1, -.5, 0, .5, 1], np.float32), 4, False, True)
    assert q.tolist() == [1, 4, 8, 12, 15] and float(s) == 7.0
    q, s, _ = O.base_quantize(np.arange(1, 10, dtype=np.float32).reshape(3, 3), 8, True, True)
    assert q.tolist() == [[146, 160, 170], [201, 207, 213], [255, 255, 255]]
    assert np.allclose(s, [[18.142859, 15.875, 14.111112]])
