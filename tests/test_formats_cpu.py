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


def test_writer_casts_parameters_to_what_the_loader_reads(tmp_path):
    """Python floats (what QuantizationState.load_state hands back), float64 and float16 / bfloat16
    scales must load back as the same values — the loader always reads float32."""
    q = torch.arange(12, dtype=torch.uint8).reshape(3, 4)
    for k, (s, z) in enumerate([(0.05, -1.25), (torch.tensor(0.05, dtype=torch.float64), torch.tensor(-1.25, dtype=torch.float64)),
                                (torch.tensor(0.5, dtype=torch.float16), torch.tensor(-1.25, dtype=torch.float16)),
                                (torch.tensor(0.5, dtype=torch.bfloat16), torch.tensor(-1.25, dtype=torch.bfloat16)),
                                (np.float64(0.05), np.float32(-1.25))]):
        path = save_quantized_tensor(q, s, z, {"bits": 8, "scheme": "asymmetric", "type": "linear"}, str(tmp_path / f"t{k}"))
        q2, s2, z2, _ = load_quantized_tensor(path)
        assert torch.equal(q2, q)
        assert abs(float(s2) - float(s)) < 1e-6 and float(z2) == -1.25
    # per-block arrays in float64 too
    s = torch.rand(3, dtype=torch.float64)
    path = save_quantized_tensor(q, s, s + 1, {"bits": 8, "blocksize": 4}, str(tmp_path / "arr"))
    _, s2, z2, _ = load_quantized_tensor(path)
    assert s2.dtype == torch.float32 and torch.allclose(s2.double(), s, atol=1e-7) and torch.allclose(z2.double(), s + 1, atol=1e-7)


def test_quantized_linear_state_dict_round_trip_and_casts():
    """a quantized layer reloads through state_dict into a fresh module; module.half() keeps the
    quantization parameters in float32 (the kernels read them as float32)."""
    from quanta_b200.nn import Linear4bit, Linear8bitLt
    for cls, kw in ((Linear4bit, {"quant_type": "linear"}), (Linear4bit, {}), (Linear8bitLt, {})):
        a = cls(128, 64, **kw)
        nbytes = 64 * 128 // 2 if a.bits == 4 else 64 * 128
        a.qweight = torch.randint(0, 255, (nbytes,), dtype=torch.uint8)
        a.scale = torch.rand(64 * 2)
        a.zero_point = None if kw == {} and cls is Linear4bit else torch.rand(64 * 2)
        a.weight = torch.nn.Parameter(torch.empty(0), requires_grad=False)
        b = cls(128, 64, **kw)
        res = b.load_state_dict(a.state_dict())
        assert not res.missing_keys and not res.unexpected_keys
        assert torch.equal(b.qweight, a.qweight) and torch.equal(b.scale, a.scale) and b.weight.numel() == 0
        assert torch.equal(b.bias, a.bias)
        h = b.half()
        assert h.scale.dtype == torch.float32 and h.bias.dtype == torch.float16 and h.qweight.dtype == torch.uint8
        if b.zero_point is not None:
            assert h.zero_point.dtype == torch.float32
    # a float checkpoint still loads into a fresh module
    c, d = Linear8bitLt(32, 16), Linear8bitLt(32, 16)
    d.load_state_dict(c.state_dict())
    assert torch.equal(c.weight, d.weight)


def test_gemm_wrappers_refuse_miscast_parameters():
    """the wrappers validate dtype / size before handing raw pointers to the kernels (checked before the
    device check would matter: these raise on any device)."""
    import pytest
    from quanta_b200.nn.functional import _check_weight_args
    x = torch.zeros(2, 128, dtype=torch.float16)
    wq = torch.zeros(64 * 128 // 2, dtype=torch.uint8)
    s = torch.zeros(64 * 2)
    _check_weight_args(x, wq, (("scale", s),), 64, 128, 4, 64, "t")
    with pytest.raises(TypeError):
        _check_weight_args(x, wq, (("scale", s.half()),), 64, 128, 4, 64, "t")
    with pytest.raises(ValueError):
        _check_weight_args(x, wq, (("scale", s[:-1]),), 64, 128, 4, 64, "t")
    with pytest.raises(ValueError):
        _check_weight_args(x, wq[:-1], (("scale", s),), 64, 128, 4, 64, "t")
    with pytest.raises(TypeError):
        _check_weight_args(x, wq.to(torch.int8), (("scale", s),), 64, 128, 4, 64, "t")
