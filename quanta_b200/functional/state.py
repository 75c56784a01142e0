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


def _jsonable(value):
    """Tensors leave as nested lists (0-dim tensors as numbers), everything else unchanged."""
    return value.detach().cpu().tolist() if isinstance(value, torch.Tensor) else value


def _from_json(value):
    """Lists come back as tensors; numbers stay numbers (so a 0-dim scale reloads as a Python float, as it
    does in the reference)."""
    return torch.tensor(value) if isinstance(value, list) else value


class QuantizationState:
    """Per-tensor / per-layer parameter tables plus a global default configuration (state.py:19-27).
    ``tensor_params``, ``layers_params`` and ``global_config`` are plain dicts, as in the reference."""

    _DEFAULTS = (("default_bits", 8), ("default_scheme", QuantizationScheme.SYMMETRIC.value),
                 ("default_type", QuantizationType.LINEAR.value))

    def __init__(self):
        self.tensor_params: Dict[str, Dict] = {}
        self.layers_params: Dict[str, Dict] = {}
        self.global_config = dict(self._DEFAULTS)
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

    # -- JSON (state.py:84-133) ----------------------------------------------------------------
    def save_state(self, filepath: str):
        document = {"tensor_params": {name: {k: _jsonable(v) for k, v in table.items()}
                                      for name, table in self.tensor_params.items()},
                    "layers_params": self.layers_params,
                    "global_config": self.global_config}
        with open(filepath, "w") as out:
            json.dump(document, out, indent=2)

    def load_state(self, filepath: str):
        if not os.path.exists(filepath):
            raise FileNotFoundError(f"State file not found: {filepath}")
        with open(filepath, "r") as src:
            document = json.load(src)
        self.global_config = document.get("global_config", self.global_config)
        self.layers_params = document.get("layers_params", {})
        for name, table in document.get("tensor_params", {}).items():
            self.tensor_params[name] = {k: _from_json(v) for k, v in table.items()}

    # -- files (state.py:135-196) --------------------------------------------------------------
    def _require(self, tensor_name: str, need_scale: bool = False) -> Dict:
        table = self.tensor_params.get(tensor_name)
        if table is None:
            raise ValueError(f"No parameters found for tensor '{tensor_name}' in state")
        if need_scale and (table.get("scale") is None or table.get("zero_point") is None):
            raise ValueError(f"Missing scale or zero_point for tensor '{tensor_name}'")
        return table

    def save_quantized_tensor_with_state(self, tensor_name: str, q_tensor: torch.Tensor, file_path: str):
        from ..utils import utils as io
        table = self._require(tensor_name, need_scale=True)
        write = io.save_quantized_tensor_torch if file_path.endswith(".pt") else io.save_quantized_tensor
        write(q_tensor, table["scale"], table["zero_point"], table, file_path)

    def load_quantized_tensor_with_state(self, tensor_name: str, file_path: str, device=None) -> torch.Tensor:
        from ..utils import utils as io
        if file_path.endswith(".pt"):
            q, scale, zero_point, table = io.load_quantized_tensor_torch(file_path, map_location=device)
        else:
            q, scale, zero_point, table = io.load_quantized_tensor(file_path, device=device)
        table.setdefault("scale", scale)
        table.setdefault("zero_point", zero_point)
        self.tensor_params[tensor_name] = table
        self._quantized_tensors[tensor_name] = q
        return q

    # -- conversions (state.py:198-286) --------------------------------------------------------
    def convert_tensor_precision(self, tensor_name: str, target_bits: int, target_type: str = "linear",
                                 target_scheme: str = None) -> torch.Tensor:
        from ..utils.utils import convert_precision
        source = self._require(tensor_name)
        held = self._quantized_tensors.get(tensor_name)
        if held is None:
            raise ValueError(f"Quantized tensor '{tensor_name}' not found in state. "
                             f"Please load the tensor first using load_quantized_tensor_with_state.")
        q, _, _, table = convert_precision(held, source, target_bits, target_type, target_scheme)
        self.tensor_params[tensor_name] = table
        self._quantized_tensors[tensor_name] = q
        return q

    def dequantize_tensor(self, tensor_name: str, q_tensor: torch.Tensor) -> torch.Tensor:
        from . import quantization as fq
        table = self._require(tensor_name, need_scale=True)
        bits = table.get("bits", 8)
        try:
            fn = {8: fq.dequantize_8bit, 4: fq.dequantize_4bit}[bits]
        except KeyError:
            raise ValueError(f"Unsupported bit depth: {bits}") from None
        return fn(q_tensor, table["scale"], table["zero_point"], quant_type=table.get("type", QuantizationType.LINEAR.value))
