"""Generate tests/golden/quanta_golden_n2.npz for the model-level sweep (row N2) with the UNMODIFIED reference.

``Quanta/functional/model.py`` imports ``onnx`` (absent here) — it is imported with an empty stand-in module in
``sys.modules``; no reference source is changed.  ``ModelQuantize._quantize_tensor`` itself raises TypeError in the
reference (it passes ``symmetric=`` to functions that do not take it, model.py:67,69 — recorded below as
``reference_quantize_tensor_raises``), so the golden values are what that call computes once the stray keyword is
dropped: ``quantize_{8,4}bit(param, quant_type, per_channel=True)`` for matrices (per tensor for 1-D parameters,
where per_channel raises), packed by the reference's own ``ModelQuantize._pack_tensor`` (first element -> high nibble).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_n2.py"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("QUANTA_REFERENCE", "/root/reference"))
sys.modules.setdefault("onnx", types.ModuleType("onnx"))
from Quanta.functional.model import ModelQuantize  # noqa: E402
from Quanta.functional.quantization import quantize_8bit, quantize_4bit  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quanta_golden_n2.npz")
torch.manual_seed(20261018)


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(96, 64)
        self.act = nn.ReLU()
        self.fc2 = nn.Linear(64, 33, bias=True)
        self.norm = nn.LayerNorm(33)

    def forward(self, x):
        return self.norm(self.fc2(self.act(self.fc1(x))))


net = Net()
store, manifest = {}, []
raises = False
try:
    ModelQuantize(net)._quantize_tensor(net.fc1.weight.data, {"bits": 8, "scheme": "symmetric", "quant_type": "linear"})
except TypeError:
    raises = True
for name, p in net.named_parameters():
    store[f"param/{name}"] = p.detach().numpy().copy()
    for bits in (8, 4):
        fn = quantize_8bit if bits == 8 else quantize_4bit
        q, s, z = fn(p.data, "linear", p.dim() > 1)
        codes = ModelQuantize._pack_tensor(None, q.reshape(-1), bits) if bits == 4 else q
        key = f"{name}/{bits}"
        manifest.append({"param": name, "bits": bits})
        store[key + "/codes"] = codes.numpy().copy()
        store[key + "/scale"] = s.numpy().copy()
        store[key + "/zp"] = z.numpy().copy()
# the nibble order itself, on the SURVEY Appendix B vector
store["pack_hi_arange16"] = ModelQuantize._pack_tensor(None, torch.arange(16, dtype=torch.uint8), 4).numpy()
store["pack_hi_odd"] = ModelQuantize._pack_tensor(None, torch.tensor([1, 2, 3], dtype=torch.uint8), 4).numpy()
store["unpack_hi_arange16"] = ModelQuantize._unpack_tensor(None, torch.from_numpy(store["pack_hi_arange16"]), (16,), 4).numpy()
store["manifest"] = np.frombuffer(json.dumps({"cases": manifest, "reference_quantize_tensor_raises": raises}).encode(), dtype=np.uint8)
np.savez_compressed(OUT, **store)
print("wrote", OUT, len(manifest), "cases; reference _quantize_tensor raises:", raises, store["pack_hi_arange16"], store["pack_hi_odd"])
