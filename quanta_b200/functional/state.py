"""Quantization bookkeeping either side of the hot path (row N3): which parameters belong to which tensor,
a JSON round trip of that table, and save / load / convert / dequantize by tensor name.  Mirrors the public
surface of Quanta/functional/state.py:7-286 (enums :7-17, QuantizationState :19-286); the tensor work is done
by the CUDA entry points of this package."""
from __future__ import annotations

import json
import os
from enum import Enum
from typing import Dict, Optional

import torch


class QuantizationScheme(Enum):
    SYMMETRIC = "symmetric"
    ASYMMETRIC = "asymmetric"


class QuantizationType(Enum):
    LINEAR = "linear"
    NF4 = "nf4"
    NF8 = "nf8"
    FP4 = "fp4"
    FP8 = "fp8"


class QuantizationState:
    """Per-tensor / per-layer parameter tables plus a global default configuration (state.py:19-27)."""

    def __init__(self):
        self.tensor_params: Dict[str, Dict] = {}
        self.layers_params: Dict[str, Dict] = {}
        self.global_config = {"default_bits": 8, "default_scheme": QuantizationScheme.SYMMETRIC.value,
                              "default_type": QuantizationType.LINEAR.value}
        self._quantized_tensors: Dict[str, torch.Tensor] = {}

    # -- tables (state.py:29-82) --------------------------------------------------------------
    def set_tensor_params(self, tensor_id: str, params: Dict):
        self.tensor_params[tensor_id] = params

    def get_tensor_params(self, tensor_id: str) -> Optional[Dict]:
        return self.tensor_params.get(tensor_id)

    def set_layer_params(self, layer_name: str, params: Dict):
        self.layers_params[layer_name] = params

    def get_layer_params(self, layer_name: str) -> Optional[Dict]:
        return self.layers_params.get(layer_name)

    def update_global_config(self, config_updates: Dict):
        self.global_config.update(config_updates)

    # -- JSON (state.py:84-133): tensors become nested lists on the way out, lists become tensors on the way in
    def save_state(self, filepath: str):
        tensors = {name: {k: (v.detach().cpu().tolist() if isinstance(v, torch.Tensor) else v) for k, v in params.items()}
                   for name, params in self.tensor_params.items()}
        with open(filepath, "w") as f:
            json.dump({"tensor_params": tensors, "layers_params": self.layers_params, "global_config": self.global_config},
                      f, indent=2)

    def load_state(self, filepath: str):
        if not os.path.exists(filepath):
            raise FileNotFoundError(f"State file not found: {filepath}")
        with open(filepath, "r") as f:
            state = json.load(f)
        self.global_config = state.get("global_config", self.global_config)
        self.layers_params = state.get("layers_params", {})
        for name, params in state.get("tensor_params", {}).items():
            self.tensor_params[name] = {k: (torch.tensor(v) if isinstance(v, list) else v) for k, v in params.items()}

    # -- files (state.py:135-196) --------------------------------------------------------------
    def _params_or_raise(self, tensor_name: str) -> Dict:
        params = self.get_tensor_params(tensor_name)
        if params is None:
            raise ValueError(f"No parameters found for tensor '{tensor_name}' in state")
        return params

    def save_quantized_tensor_with_state(self, tensor_name: str, q_tensor: torch.Tensor, file_path: str):
        from ..utils.utils import save_quantized_tensor, save_quantized_tensor_torch
        params = self._params_or_raise(tensor_name)
        scale, zero_point = params.get("scale"), params.get("zero_point")
        if scale is None or zero_point is None:
            raise ValueError(f"Missing scale or zero_point for tensor '{tensor_name}'")
        writer = save_quantized_tensor_torch if file_path.endswith(".pt") else save_quantized_tensor
        writer(q_tensor, scale, zero_point, params, file_path)

    def load_quantized_tensor_with_state(self, tensor_name: str, file_path: str, device=None) -> torch.Tensor:
        from ..utils.utils import load_quantized_tensor, load_quantized_tensor_torch
        if file_path.endswith(".pt"):
            q, scale, zero_point, params = load_quantized_tensor_torch(file_path, map_location=device)
        else:
            q, scale, zero_point, params = load_quantized_tensor(file_path, device=device)
        params.setdefault("scale", scale)
        params.setdefault("zero_point", zero_point)
        self.set_tensor_params(tensor_name, params)
        self._quantized_tensors[tensor_name] = q
        return q

    # -- conversions (state.py:198-286) --------------------------------------------------------
    def convert_tensor_precision(self, tensor_name: str, target_bits: int, target_type: str = "linear",
                                 target_scheme: str = None) -> torch.Tensor:
        from ..utils.utils import convert_precision
        source = self._params_or_raise(tensor_name)
        if tensor_name not in self._quantized_tensors:
            raise ValueError(f"Quantized tensor '{tensor_name}' not found in state. "
                             f"Please load the tensor first using load_quantized_tensor_with_state.")
        q, _, _, new_params = convert_precision(self._quantized_tensors[tensor_name], source, target_bits, target_type,
                                                target_scheme)
        self.set_tensor_params(tensor_name, new_params)
        self._quantized_tensors[tensor_name] = q
        return q

    def dequantize_tensor(self, tensor_name: str, q_tensor: torch.Tensor) -> torch.Tensor:
        from .quantization import dequantize_4bit, dequantize_8bit
        params = self._params_or_raise(tensor_name)
        bits = params.get("bits", 8)
        scale, zero_point = params.get("scale"), params.get("zero_point")
        if scale is None or zero_point is None:
            raise ValueError(f"Missing scale or zero_point for tensor '{tensor_name}'")
        if bits not in (4, 8):
            raise ValueError(f"Unsupported bit depth: {bits}")
        fn = dequantize_8bit if bits == 8 else dequantize_4bit
        return fn(q_tensor, scale, zero_point, quant_type=params.get("type", QuantizationType.LINEAR.value))
