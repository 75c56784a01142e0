"""Golden artefacts for row N3 produced by the UNMODIFIED reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_n3.py

  ref_tensor8.qtn / ref_tensor4.qtn   files written by Quanta.utils.utils.save_quantized_tensor
  ref_state.json                      QuantizationState.save_state output
  quanta_golden_n3.npz                inputs, what the reference loader returns, convert_precision results"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("QUANTA_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from Quanta.functional.quantization import quantize_8bit, quantize_4bit  # noqa: E402
from Quanta.functional.state import QuantizationState  # noqa: E402
from Quanta.utils.utils import (save_quantized_tensor, load_quantized_tensor, convert_precision)  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(1357)
x = torch.randn(24, 40, generator=g) * 0.05
store = {"x": x.numpy()}

q8, s8, z8 = quantize_8bit(x)
q4, s4, z4 = quantize_4bit(x)
for name, (q, s, z, bits) in {"ref_tensor8": (q8, s8, z8, 8), "ref_tensor4": (q4, s4, z4, 4)}.items():
    path = os.path.join(HERE, name + ".qtn")
    save_quantized_tensor(q, s, z, {"bits": bits, "scheme": "asymmetric", "type": "linear"}, path)
    lq, ls, lz, meta = load_quantized_tensor(path)
    store[f"{name}/q"], store[f"{name}/scale"], store[f"{name}/zp"] = lq.numpy(), ls.numpy(), lz.numpy()

# convert_precision: 8 -> 4 and 4 -> 8 (linear), 8 -> nf4, 8 -> fp4, 4 -> nf8, 4 -> fp8
src8 = {"bits": 8, "type": "linear", "scheme": "asymmetric", "scale": s8, "zero_point": z8}
src4 = {"bits": 4, "type": "linear", "scheme": "asymmetric", "scale": s4, "zero_point": z4}
for tag, (q, src, bits, typ) in {"c84_linear": (q8, src8, 4, "linear"), "c48_linear": (q4, src4, 8, "linear"),
                                 "c84_nf4": (q8, src8, 4, "nf4"), "c84_fp4": (q8, src8, 4, "fp4"),
                                 "c48_nf8": (q4, src4, 8, "nf8"), "c48_fp8": (q4, src4, 8, "fp8")}.items():
    nq, ns, nz, _ = convert_precision(q, src, bits, typ)
    store[f"{tag}/q"] = nq.numpy()
    store[f"{tag}/a"] = np.zeros(0, np.float32) if ns is None else ns.numpy()
    store[f"{tag}/b"] = np.asarray(nz.numpy() if isinstance(nz, torch.Tensor) else nz, dtype=np.float32)

st = QuantizationState()
st.set_tensor_params("w", {"bits": 8, "type": "linear", "scheme": "asymmetric", "scale": s8, "zero_point": z8})
st.set_layer_params("fc1", {"bits": 4, "type": "nf4"})
st.update_global_config({"default_bits": 4})
st.save_state(os.path.join(HERE, "ref_state.json"))
np.savez_compressed(os.path.join(HERE, "quanta_golden_n3.npz"), **store)
print("ok", sorted(store)[:6])
