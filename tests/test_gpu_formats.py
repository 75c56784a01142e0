"""Row N3 on the GPU: convert_precision (dequantize + requantize, both on the device) reproduces the
reference's results for 8 <-> 4 bit and every target type; files round-trip through CUDA tensors; the
QuantizationState helpers drive the same kernels."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Z = np.load(os.path.join(G, "quanta_golden_n3.npz"))


def _src(name, bits):
    from quanta_b200.utils import load_quantized_tensor
    q, s, z, _ = load_quantized_tensor(os.path.join(G, name + ".qtn"), device="cuda")
    return q, {"bits": bits, "type": "linear", "scheme": "asymmetric", "scale": s, "zero_point": z}


@pytest.mark.parametrize("tag,src,bits,typ", [("c84_linear", "ref_tensor8", 4, "linear"), ("c48_linear", "ref_tensor4", 8, "linear"),
                                              ("c84_nf4", "ref_tensor8", 4, "nf4"), ("c84_fp4", "ref_tensor8", 4, "fp4"),
                                              ("c48_nf8", "ref_tensor4", 8, "nf8"), ("c48_fp8", "ref_tensor4", 8, "fp8")])
def test_convert_precision_matches_reference(tag, src, bits, typ):
    from quanta_b200.utils import convert_precision
    q, params = _src(src, 8 if src.endswith("8") else 4)
    nq, a, b, new_params = convert_precision(q, params, bits, typ)
    assert nq.is_cuda and np.array_equal(nq.cpu().numpy(), Z[f"{tag}/q"])
    if a is None:
        assert Z[f"{tag}/a"].size == 0
    else:
        assert np.array_equal(a.cpu().numpy().view(np.uint32), Z[f"{tag}/a"].view(np.uint32))
    bb = b.cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float32)
    assert np.array_equal(np.asarray(bb, np.float32).view(np.uint32), Z[f"{tag}/b"].view(np.uint32))
    assert new_params["bits"] == bits and new_params["type"] == typ and new_params["shape"] == tuple(nq.shape)


def test_state_round_trip_on_cuda(tmp_path):
    import quanta_b200 as Q
    from quanta_b200.functional.state import QuantizationState
    x = torch.from_numpy(Z["x"]).cuda()
    q, s, z = Q.quantize_8bit(x)
    st = QuantizationState()
    st.set_tensor_params("w", {"bits": 8, "type": "linear", "scheme": "asymmetric", "scale": s, "zero_point": z})
    path = str(tmp_path / "w.qtn")
    st.save_quantized_tensor_with_state("w", q, path)
    assert open(path, "rb").read() == open(os.path.join(G, "ref_tensor8.qtn"), "rb").read()      # same bytes as the reference
    st2 = QuantizationState()
    q2 = st2.load_quantized_tensor_with_state("w", path, device="cuda")
    assert torch.equal(q2, q)
    d = st2.dequantize_tensor("w", q2)
    assert torch.equal(d, Q.dequantize_8bit(q, s, z))
    q4 = st2.convert_tensor_precision("w", 4)
    assert np.array_equal(q4.cpu().numpy(), Z["c84_linear/q"]) and st2.get_tensor_params("w")["bits"] == 4


@pytest.mark.parametrize("n", [1, 17, 4096, 1000003])
@pytest.mark.parametrize("src_bits,tgt_bits", [(8, 4), (4, 8), (8, 8), (4, 4)])
def test_fused_convert_equals_dequantize_plus_quantize(n, src_bits, tgt_bits):
    """quanta_convert_linear (a code -> code table, 3 B/element) against the two-step path it replaces, bit for bit,
    incl. sparse code sets (min / max over the codes that occur), negative and non-finite scales."""
    import quanta_b200 as Q
    from quanta_b200.utils import convert_precision
    g = torch.Generator().manual_seed(n + src_bits)
    top = 256 if src_bits == 8 else 16
    for trial, (s, z) in enumerate([(0.0123, -1.5), (3.7e-5, 0.25), (-0.5, 2.0), (float("inf"), 0.0), (0.0, 7.0), (1e30, -1e30)]):
        q = torch.randint(3 if trial % 2 else 0, top - (5 if trial % 2 else 0), (n,), dtype=torch.uint8, generator=g).cuda()
        params = {"bits": src_bits, "type": "linear", "scheme": "asymmetric", "scale": torch.tensor(s), "zero_point": torch.tensor(z)}
        nq, ns, nz, newp = convert_precision(q, params, tgt_bits, "linear")
        deq = (Q.dequantize_8bit if src_bits == 8 else Q.dequantize_4bit)(q, torch.tensor(s).cuda(), torch.tensor(z).cuda())
        rq, rs, rz = (Q.quantize_8bit if tgt_bits == 8 else Q.quantize_4bit)(deq)
        assert torch.equal(nq, rq), (trial, s, z)
        for a, b in ((ns, rs), (nz, rz)):
            a, b = a.reshape(()).cpu(), b.reshape(()).cpu()
            assert (torch.isnan(a) and torch.isnan(b)) or a.view(torch.int32) == b.view(torch.int32), (trial, s, z)
        assert newp["bits"] == tgt_bits and newp["shape"] == tuple(q.shape)
