"""Row N3 on the CPU: the .qtn writer is byte-identical to the reference's, the loader returns what the
reference loader returns for reference-written files, parameter arrays round-trip, and QuantizationState
reads and writes the reference's JSON (golden files from tests/golden/make_golden_n3.py)."""
import json
import os

import numpy as np
import torch

from quanta_b200.functional.state import QuantizationState
from quanta_b200.utils import (save_quantized_tensor, load_quantized_tensor, save_quantized_tensor_torch,
                               load_quantized_tensor_torch)

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Z = np.load(os.path.join(G, "quanta_golden_n3.npz"))


def test_loader_reads_reference_files_like_the_reference():
    for name, bits in (("ref_tensor8", 8), ("ref_tensor4", 4)):
        q, s, z, meta = load_quantized_tensor(os.path.join(G, name + ".qtn"))
        assert q.dtype == torch.uint8 and np.array_equal(q.numpy(), Z[f"{name}/q"])
        assert s.dim() == 0 and s.numpy().tobytes() == Z[f"{name}/scale"].tobytes()
        assert z.numpy().tobytes() == Z[f"{name}/zp"].tobytes()
        assert meta["bits"] == bits and meta["scheme"] == "asymmetric" and meta["type"] == "linear"


def test_writer_is_byte_identical_to_the_reference(tmp_path):
    for name, bits in (("ref_tensor8", 8), ("ref_tensor4", 4)):
        q = torch.from_numpy(Z[f"{name}/q"])
        s, z = torch.tensor(Z[f"{name}/scale"].item()), torch.tensor(Z[f"{name}/zp"].item())
        out = save_quantized_tensor(q, s, z, {"bits": bits, "scheme": "asymmetric", "type": "linear"}, str(tmp_path / name))
        assert out.endswith(".qtn")
        assert open(out, "rb").read() == open(os.path.join(G, name + ".qtn"), "rb").read()


def test_parameter_arrays_round_trip(tmp_path):
    q = torch.randint(0, 16, (8, 64), dtype=torch.uint8)
    s, z = torch.rand(8) + 0.1, torch.randn(8)
    path = save_quantized_tensor(q, s, z, {"bits": 4, "type": "linear", "blocksize": 64}, str(tmp_path / "blk"))
    q2, s2, z2, meta = load_quantized_tensor(path)
    assert torch.equal(q, q2) and torch.equal(s, s2) and torch.equal(z, z2) and meta["blocksize"] == 64
    p2 = save_quantized_tensor_torch(q, s, z, {"bits": 4}, str(tmp_path / "blk"))
    q3, s3, z3, params = load_quantized_tensor_torch(p2)
    assert p2.endswith(".pt") and torch.equal(q, q3) and torch.equal(s, s3) and params == {"bits": 4}


def test_state_reads_and_writes_the_reference_json(tmp_path):
    st = QuantizationState()
    st.load_state(os.path.join(G, "ref_state.json"))
    assert st.global_config["default_bits"] == 4 and st.get_layer_params("fc1") == {"bits": 4, "type": "nf4"}
    p = st.get_tensor_params("w")
    # a 0-dim tensor is written as a JSON number and (like in the reference) comes back as a Python float
    assert p["bits"] == 8 and isinstance(p["scale"], float)
    assert np.float32(p["scale"]).tobytes() == Z["ref_tensor8/scale"].tobytes()
    out = tmp_path / "state.json"
    st.save_state(str(out))
    assert json.load(open(out)) == json.load(open(os.path.join(G, "ref_state.json")))
    assert open(out).read() == open(os.path.join(G, "ref_state.json")).read()
