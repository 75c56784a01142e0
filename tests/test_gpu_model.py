"""Row N2 on the GPU: ``quanta_b200.functional.model.ModelQuantize`` against golden values computed by the
unmodified reference (tests/golden/make_golden_n2.py): per-parameter per_channel=True codes / scale / zero-point,
the sweep's high-nibble-first packing, the recorded QuantizationState, and the batched blockwise variant."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quanta_golden_n2.npz"))
META = json.loads(bytes(Z["manifest"]).decode())


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(96, 64)
        self.act = nn.ReLU()
        self.fc2 = nn.Linear(64, 33, bias=True)
        self.norm = nn.LayerNorm(33)

    def forward(self, x):
        return self.norm(self.fc2(self.act(self.fc1(x))))


def golden_net():
    net = Net()
    with torch.no_grad():
        for name, p in net.named_parameters():
            p.copy_(torch.from_numpy(Z[f"param/{name}"]))
    return net.cuda()


def bits_equal(t, ref):
    a = t.detach().cpu().numpy().astype(np.float32)
    return a.shape == np.asarray(ref).shape and np.array_equal(a.view(np.uint32), np.asarray(ref, np.float32).view(np.uint32))


def test_nibble_order_of_the_sweep():
    from quanta_b200.functional.model import ModelQuantize
    mq = ModelQuantize(nn.Identity())
    assert META["reference_quantize_tensor_raises"] is True       # why the golden is what the call computes without `symmetric=`
    a = torch.arange(16, dtype=torch.uint8, device="cuda")
    assert np.array_equal(mq._pack_tensor(a, 4).cpu().numpy(), Z["pack_hi_arange16"])
    assert np.array_equal(mq._pack_tensor(torch.tensor([1, 2, 3], dtype=torch.uint8, device="cuda"), 4).cpu().numpy(), Z["pack_hi_odd"])
    assert np.array_equal(mq._unpack_tensor(torch.from_numpy(Z["pack_hi_arange16"]).cuda(), (16,), 4).cpu().numpy(), Z["unpack_hi_arange16"])
    g = torch.Generator().manual_seed(1)
    for n in (1, 2, 15, 16, 17, 4097, 100000):
        q = torch.randint(0, 16, (n,), dtype=torch.uint8, generator=g)
        ref = torch.cat([q, torch.zeros(n % 2, dtype=torch.uint8)]).reshape(-1, 2)
        ref = ((ref[:, 0] << 4) | ref[:, 1]).numpy()
        pk = mq._pack_tensor(q.cuda(), 4)
        assert np.array_equal(pk.cpu().numpy(), ref)
        assert torch.equal(mq._unpack_tensor(pk, (n,), 4).cpu(), q)
    assert mq._pack_tensor(a, 8) is a


@pytest.mark.parametrize("bits", [8, 4])
def test_sweep_matches_reference_per_parameter(bits):
    from quanta_b200.functional.model import ModelQuantize
    net = golden_net()
    mq = ModelQuantize(net, bits=bits, scheme="asymmetric")
    qm = mq.quantize()
    assert qm is not net and net.fc1.weight.dtype == torch.float32          # the original model is untouched
    names = [n for n, _ in net.named_parameters()]
    assert sorted(mq.state.tensor_params) == sorted(names)
    for name, p in qm.named_parameters():
        rec = mq.state.get_tensor_params(name)
        assert rec["bits"] == bits and rec["scheme"] == "asymmetric" and rec["quant_type"] == "linear"
        assert tuple(rec["original_shape"]) == tuple(Z[f"param/{name}"].shape)
        assert p.dtype == torch.uint8 and not p.requires_grad
        assert np.array_equal(p.data.cpu().numpy(), Z[f"{name}/{bits}/codes"]), name
        assert bits_equal(rec["scale"], Z[f"{name}/{bits}/scale"]) and bits_equal(rec["zero_point"], Z[f"{name}/{bits}/zp"]), name
        d = mq.dequantize_parameter(name, p.data)
        step = float(rec["scale"].max())
        assert float((d.cpu() - torch.from_numpy(Z[f"param/{name}"])).abs().max()) <= 0.51 * step + 1e-7


def test_layer_config_and_errors():
    from quanta_b200.functional.model import ModelQuantize
    net = golden_net()
    mq = ModelQuantize(net, bits=8)
    mq.config_layer("fc2", bits=4, scheme="asymmetric")
    assert mq._get_layer_config("fc1")["bits"] == 8 and mq._get_layer_config("fc2") == {
        "bits": 4, "scheme": "asymmetric", "weights_only": False, "quant_type": "linear", "calibration_method": "minmax"}
    qm = mq.quantize()
    assert qm.fc1.weight.shape == (64, 96) and qm.fc2.weight.numel() == (33 * 64 + 1) // 2
    assert np.array_equal(qm.fc2.weight.data.cpu().numpy(), Z["fc2.weight/4/codes"])
    assert np.array_equal(qm.fc1.weight.data.cpu().numpy(), Z["fc1.weight/8/codes"])
    with pytest.raises(ValueError):
        ModelQuantize(net, bits=3).quantize()


@pytest.mark.parametrize("bits", [8, 4])
def test_blockwise_batched_sweep(bits):
    """blocksize=64: the multi-tensor launch; same codes as one quantize_*bit(blocksize=64) call per parameter."""
    import quanta_b200 as Q
    from quanta_b200.functional.model import ModelQuantize
    torch.manual_seed(3)
    net = nn.Sequential(*[nn.Linear(256, 256, bias=True) for _ in range(20)]).cuda()
    mq = ModelQuantize(net, bits=bits, blocksize=64)
    qm = mq.quantize()
    for (name, p), (_, q) in zip(net.named_parameters(), qm.named_parameters()):
        fn = Q.quantize_4bit if bits == 4 else Q.quantize_8bit
        qr, sr, zr = fn(p.data, blocksize=64)
        codes = mq._pack_tensor(qr, bits) if bits == 4 else qr
        rec = mq.state.get_tensor_params(name)
        assert torch.equal(q.data, codes) and torch.equal(rec["scale"], sr) and torch.equal(rec["zero_point"], zr)
        assert torch.allclose(mq.dequantize_parameter(name, q.data), p.data, atol=float(sr.max()) * 0.51)
