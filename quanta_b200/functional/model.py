"""The model-level sweep of ``Quanta.functional.model.ModelQuantize`` (Quanta/functional/model.py:8-118, :254-289;
row N2): walk a module tree, quantize every trainable parameter ``per_channel=True``, pack 4-bit codes with the
sweep's own nibble order (first element -> HIGH nibble, model.py:73-82 — the opposite of ``pack_4bit_tensor``),
record the parameters in a :class:`QuantizationState` and replace ``param.data`` by the codes.

The reference's class cannot run as written: it imports ``onnx`` (absent), passes ``symmetric=`` to
functions that do not take it (model.py:67,69 -> TypeError), its Python packing loop indexes a 2-D tensor like
a flat one, and ``named_modules()`` x ``named_parameters()`` visits every nested parameter once per ancestor.
What is kept is its contract — constructor, ``config_layer`` / ``_get_layer_config``, ``_quantize_tensor`` /
``_pack_tensor`` / ``_unpack_tensor``, ``quantize()`` and the state it records (keys ``bits, scheme, quant_type,
scale, zero_point, original_shape`` under ``"<module>.<param>"``); what is defined here, because the reference
leaves it broken:

* ``scheme`` is recorded but does not change the arithmetic (the public quantize_*bit functions are min-max
  affine, convention A) — exactly what the reference's call would do if the stray keyword were dropped;
* 1-D parameters (biases) are quantized per tensor (``per_channel=True`` raises on them in the reference);
* every parameter is quantized once, under the name of the module that owns it;
* codes are packed in flat row-major order.

All tensor work runs on the CUDA entry points; ``blocksize=B`` switches the sweep to blockwise quantization
through the multi-tensor launch (``quanta_quantize_block_batch``: 16 parameters per kernel) — the form the
headline benchmark measures."""
from __future__ import annotations

import copy
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import _host, _lib
from .quantization import quantize_4bit, quantize_8bit, quantize_4bit_many, quantize_8bit_many
from .state import QuantizationState


class ModelQuantize:
    def __init__(self, model: nn.Module, bits: int = 8, scheme: str = "symmetric", quant_type: str = "linear",
                 blocksize: Optional[int] = None):
        self.model = model
        self.bits = bits
        self.scheme = scheme
        self.quant_type = quant_type
        self.blocksize = blocksize
        self.layer_config: Dict[str, Dict[str, Any]] = {}
        self.quantized_model = None
        self.state = QuantizationState()

    # ---- configuration (model.py:25-58) -----------------------------------------------------------
    def config_layer(self, layer_name: str, bits: int, scheme: str = None, weights_only: bool = False,
                     quant_type: str = None, calibration_method: str = "minmax"):
        self.layer_config[layer_name] = {
            "bits": bits,
            "scheme": scheme if scheme is not None else self.scheme,
            "weights_only": weights_only,
            "quant_type": quant_type if quant_type is not None else self.quant_type,
            "calibration_method": calibration_method,
        }

    def _get_layer_config(self, layer_name: str) -> Dict[str, Any]:
        return self.layer_config.get(layer_name, {
            "bits": self.bits, "scheme": self.scheme, "weights_only": False, "quant_type": self.quant_type,
            "calibration_method": "minmax"})

    # ---- per-tensor pieces (model.py:60-94) ---------------------------------------------------------
    def _quantize_tensor(self, tensor: torch.Tensor, config: Dict[str, Any]) -> tuple:
        bits, quant_type = config["bits"], config["quant_type"]
        if bits not in (4, 8):
            raise ValueError(f"Unsupported bit depth: {bits}")
        fn = quantize_8bit if bits == 8 else quantize_4bit
        if self.blocksize is not None and quant_type == "linear":
            return fn(tensor, quant_type=quant_type, blocksize=self.blocksize)
        return fn(tensor, quant_type=quant_type, per_channel=tensor.dim() > 1)

    def _pack_tensor(self, tensor: torch.Tensor, bits: int) -> torch.Tensor:
        """4-bit: ``packed[i] = (t[2i] << 4) | t[2i+1]`` over the flattened codes, one zero pad if odd."""
        if bits != 4:
            return tensor
        _host.require_cuda(tensor)
        q = tensor.detach().reshape(-1)
        if q.dtype != torch.uint8:
            q = q.to(torch.uint8)
        q = q.contiguous()
        out = torch.empty((q.numel() + 1) // 2, dtype=torch.uint8, device=q.device)
        if q.numel():
            with _host.device_guard(q.device):
                st = _lib.lib().quanta_pack4_hi(q.data_ptr(), q.numel(), out.data_ptr(), _host.stream_ptr(q.device))
            _lib.check(st, "quanta_pack4_hi")
        return out

    def _unpack_tensor(self, packed: torch.Tensor, original_shape: Tuple[int, ...], bits: int) -> torch.Tensor:
        if bits != 4:
            return packed
        _host.require_cuda(packed, "packed")
        p = packed.detach().reshape(-1).contiguous()
        n = 1
        for d in original_shape:
            n *= int(d)
        if p.numel() != (n + 1) // 2:
            raise ValueError(f"packed codes hold {p.numel()} bytes, shape {tuple(original_shape)} needs {(n + 1) // 2}")
        out = torch.empty(p.numel() * 2, dtype=torch.uint8, device=p.device)
        if p.numel():
            with _host.device_guard(p.device):
                st = _lib.lib().quanta_unpack4_hi(p.data_ptr(), p.numel(), out.data_ptr(), _host.stream_ptr(p.device))
            _lib.check(st, "quanta_unpack4_hi")
        return out[:n].reshape(tuple(original_shape))

    # ---- the sweep (model.py:96-118, :254-289) ------------------------------------------------------
    def _record(self, key, config, scale, zero_point, shape):
        self.state.set_tensor_params(key, {
            "bits": config["bits"], "scheme": config["scheme"], "quant_type": config["quant_type"],
            "scale": scale, "zero_point": zero_point, "original_shape": shape})

    def _quantize_module(self, module: nn.Module, name: str) -> None:
        """Quantize the parameters this module owns directly (children are visited under their own names)."""
        config = self._get_layer_config(name)
        for param_name, param in module._parameters.items():
            if param is None or not param.requires_grad:
                continue
            q_tensor, scale, zero_point = self._quantize_tensor(param.data, config)
            shape = param.data.shape
            if config["bits"] == 4:
                q_tensor = self._pack_tensor(q_tensor, 4)
            self._record(f"{name}.{param_name}", config, scale, zero_point, shape)
            param.requires_grad_(False)                 # integer codes carry no gradient
            param.data = q_tensor

    def quantize(self, calibration_data=None) -> nn.Module:
        """Returns a quantized deep copy of the model (the reference rebuilds it with ``type(model)()`` +
        ``load_state_dict``, model.py:263-264, which only works for argument-less constructors).
        ``calibration_data`` is accepted for signature compatibility; weight quantization does not use it."""
        self.quantized_model = copy.deepcopy(self.model)
        owners = [(n, m) for n, m in self.quantized_model.named_modules()
                  if any(p is not None and p.requires_grad for p in m._parameters.values())]
        if self.blocksize is not None:
            self._quantize_blockwise_batched(owners)
        else:
            for name, module in owners:
                self._quantize_module(module, name)
        return self.quantized_model

    def _quantize_blockwise_batched(self, owners):
        """blocksize=B: all parameters that share (bits, "linear") go through the multi-tensor launch."""
        groups = {4: [], 8: []}
        rest = []
        for name, module in owners:
            config = self._get_layer_config(name)
            for pname, param in module._parameters.items():
                if param is None or not param.requires_grad:
                    continue
                ok = (config["quant_type"] == "linear" and config["bits"] in (4, 8) and param.is_cuda
                      and param.numel() % self.blocksize == 0 and param.numel() > 0)
                (groups[config["bits"]] if ok else rest).append((name, pname, param, config))
        for bits, items in groups.items():
            if not items:
                continue
            tensors = [p.data for _, _, p, _ in items]
            if bits == 4:
                outs = quantize_4bit_many(tensors, blocksize=self.blocksize, packed=False)
            else:
                outs = quantize_8bit_many(tensors, blocksize=self.blocksize)
            for (name, pname, param, config), (q, s, z) in zip(items, outs):
                shape = param.data.shape
                if bits == 4:
                    q = self._pack_tensor(q, 4)
                self._record(f"{name}.{pname}", config, s, z, shape)
                param.requires_grad_(False)
                param.data = q
        for name, pname, param, config in rest:
            q, s, z = self._quantize_tensor(param.data, config)
            shape = param.data.shape
            if config["bits"] == 4:
                q = self._pack_tensor(q, 4)
            self._record(f"{name}.{pname}", config, s, z, shape)
            param.requires_grad_(False)
            param.data = q

    def dequantize_parameter(self, key: str, codes: torch.Tensor) -> torch.Tensor:
        """The float32 tensor a recorded parameter stands for (unpack with the sweep's nibble order, then the
        public dequantize_*bit)."""
        from .quantization import dequantize_4bit, dequantize_8bit
        p = self.state.get_tensor_params(key)
        if p is None:
            raise KeyError(key)
        q = self._unpack_tensor(codes, tuple(p["original_shape"]), p["bits"]) if p["bits"] == 4 else codes
        q = q.reshape(tuple(p["original_shape"]))
        fn = dequantize_4bit if p["bits"] == 4 else dequantize_8bit
        blockwise = self.blocksize is not None and p["scale"].numel() == q.numel() // self.blocksize and p["scale"].dim() == 1 \
            and q.numel() % self.blocksize == 0 and p["scale"].numel() > 1
        return fn(q, p["scale"], p["zero_point"], p["quant_type"], **({"blocksize": self.blocksize} if blockwise else {}))
