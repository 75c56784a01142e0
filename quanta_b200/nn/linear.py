"""Drop-in CUDA versions of ``Quanta.nn.Linear8bitLt`` / ``Linear4bit``.

Constructor signatures and attributes are those of the reference
(Quanta/nn/linear.py:14-21, :52-59).  The reference's ``forward`` is a
placeholder ``F.linear(x, self.weight, self.bias)`` on floating-point weights
(:43-45, :81-83); here the weight is held quantized (convention A, blockwise
along in_features) and ``forward`` is the fused dequantize-then-matmul kernel:

    y = x @ dequantize_*bit(Wq, scale, zero_point).to(x.dtype).T + bias

Parity: defined by that composition; tolerance 1e-2 relative in
bf16/fp16.  ``Linear4bit`` supports ``quant_type="nf4"`` (the reference default: NF4 codebook,
blockwise abs_max) and ``quant_type="linear"`` (blockwise affine).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..functional.quantization import quantize_4bit, quantize_8bit, dequantize_4bit, dequantize_8bit
from .functional import linear_wna16, linear_nf4a16, int8_outlier_matmul, rowwise_quantize_sym


class _QuantLinearBase(nn.Module):
    bits = 8

    def _init_common(self, in_features, out_features, bias, blocksize):
        self.in_features = in_features
        self.out_features = out_features
        self.blocksize = blocksize
        # like the reference, the module is born with a floating-point weight ...
        self.weight = nn.Parameter(torch.empty((out_features, in_features)))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        # ... and holds the quantized form once quantize_() / load_float_weight() ran
        self.register_buffer("qweight", None, persistent=True)
        self.register_buffer("scale", None, persistent=True)
        self.register_buffer("zero_point", None, persistent=True)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    # ---- quantized state -------------------------------------------------
    def _check_quant_type(self):
        pass

    @torch.no_grad()
    def quantize_(self):
        """Quantize the current floating-point weight on its (CUDA) device and drop it."""
        self._check_quant_type()
        w = self.weight.detach()
        if self.in_features % self.blocksize:
            raise ValueError(f"in_features ({self.in_features}) must be a multiple of blocksize ({self.blocksize})")
        if self.bits == 4:
            q, s, z = quantize_4bit(w, blocksize=self.blocksize, packed=True)
        else:
            q, s, z = quantize_8bit(w, blocksize=self.blocksize)
        self.qweight, self.scale, self.zero_point = q, s, z
        self.weight = nn.Parameter(torch.empty(0, device=w.device), requires_grad=False)
        return self

    @torch.no_grad()
    def load_float_weight(self, weight, bias=None):
        """Install a floating-point [out, in] weight (and bias) and quantize it."""
        if tuple(weight.shape) != (self.out_features, self.in_features):
            raise ValueError("weight shape mismatch")
        self.weight = nn.Parameter(weight.detach().clone(), requires_grad=False)
        if bias is not None and self.bias is not None:
            self.bias.data = bias.detach().to(self.bias.dtype).clone()
        return self.quantize_()

    def dequantize_weight(self, dtype=torch.float32):
        """The [out, in] floating-point weight the forward pass multiplies with."""
        shape = (self.out_features, self.in_features)
        if self.bits == 4:
            return dequantize_4bit(self.qweight, self.scale, self.zero_point, blocksize=self.blocksize, packed=True,
                                   shape=shape, out_dtype=dtype)
        return dequantize_8bit(self.qweight.reshape(shape), self.scale, self.zero_point, blocksize=self.blocksize,
                               out_dtype=dtype)

    def _forward_quantized(self, x, compute_dtype):
        if self.qweight is None:
            if not self.weight.is_cuda:
                raise RuntimeError("quanta_b200 layers run on CUDA only: move the module to a GPU first")
            self.quantize_()
        xin = x if x.dtype == compute_dtype else x.to(compute_dtype)
        y = linear_wna16(xin, self.qweight, self.scale, self.zero_point, self.bias, bits=self.bits,
                         blocksize=self.blocksize, out_features=self.out_features, in_features=self.in_features)
        return y if y.dtype == x.dtype or not x.is_floating_point() else y.to(x.dtype)


class Linear8bitLt(_QuantLinearBase):
    """8-bit quantized linear layer (Quanta/nn/linear.py:10-45).

    ``outlier_split=False`` (default): blockwise W8A16 dequant-GEMM (row G2).
    ``outlier_split=True``: the LLM.int8() decomposition the reference's
    ``threshold`` argument points at (row G3) — row-wise symmetric int8 weight
    codes, int8 x int8 tensor-core product for the regular feature columns and
    a 16-bit product for the columns whose activations exceed ``threshold``."""
    bits = 8

    def __init__(self, in_features, out_features, bias=True, has_fp16_weights=False, threshold=6.0, blocksize=64,
                 outlier_split=False):
        super().__init__()
        self.threshold = threshold
        self.has_fp16_weights = has_fp16_weights
        self.outlier_split = outlier_split
        self._init_common(in_features, out_features, bias, blocksize)
        self.register_buffer("qweight_rowwise", None, persistent=True)
        self.register_buffer("row_scale", None, persistent=True)

    @torch.no_grad()
    def quantize_(self):
        if not self.outlier_split:
            return super().quantize_()
        w = self.weight.detach()
        self.qweight_rowwise, self.row_scale = rowwise_quantize_sym(w)
        self.weight = nn.Parameter(torch.empty(0, device=w.device), requires_grad=False)
        return self

    def dequantize_weight(self, dtype=torch.float32):
        if not self.outlier_split:
            return super().dequantize_weight(dtype)
        return (self.qweight_rowwise.float() / self.row_scale[:, None]).to(dtype)

    def forward(self, x):
        dt = x.dtype if x.dtype in (torch.float16, torch.bfloat16) else torch.float16
        if not self.outlier_split:
            return self._forward_quantized(x, dt)
        if self.qweight_rowwise is None:
            if not self.weight.is_cuda:
                raise RuntimeError("quanta_b200 layers run on CUDA only: move the module to a GPU first")
            self.quantize_()
        xin = x if x.dtype == dt else x.to(dt)
        y = int8_outlier_matmul(xin, self.qweight_rowwise, self.row_scale, self.threshold, self.bias)
        return y if y.dtype == x.dtype or not x.is_floating_point() else y.to(x.dtype)


class Linear4bit(_QuantLinearBase):
    """4-bit quantized linear layer (Quanta/nn/linear.py:48-83).

    ``quant_type="nf4"`` (the reference default): NF4 codebook with one abs_max per block of
    ``blocksize`` input features; ``quant_type="linear"``: blockwise affine min-max (convention A)."""
    bits = 4

    def __init__(self, in_features, out_features, bias=True, compute_dtype=torch.float16, quant_type="nf4",
                 blocksize=64):
        super().__init__()
        self.compute_dtype = compute_dtype
        self.quant_type = quant_type
        self._init_common(in_features, out_features, bias, blocksize)

    def _check_quant_type(self):
        if self.quant_type not in ("linear", "nf4"):
            raise NotImplementedError(f"quant_type={self.quant_type!r}: only 'nf4' and 'linear' have a fused dequant-GEMM; "
                                      "fp4 is available as quantize_4bit / dequantize_4bit (SURVEY §8(f) N4)")

    @torch.no_grad()
    def quantize_(self):
        if self.quant_type != "nf4":
            return super().quantize_()
        w = self.weight.detach()
        if self.in_features % self.blocksize:
            raise ValueError(f"in_features ({self.in_features}) must be a multiple of blocksize ({self.blocksize})")
        # scale holds abs_max per block; there is no zero-point
        self.qweight, _, self.scale = quantize_4bit(w, quant_type="nf4", blocksize=self.blocksize, packed=True)
        self.zero_point = None
        self.weight = nn.Parameter(torch.empty(0, device=w.device), requires_grad=False)
        return self

    def dequantize_weight(self, dtype=torch.float32):
        if self.quant_type != "nf4":
            return super().dequantize_weight(dtype)
        return dequantize_4bit(self.qweight, None, self.scale, quant_type="nf4", blocksize=self.blocksize, packed=True,
                               shape=(self.out_features, self.in_features), out_dtype=dtype)

    def forward(self, x):
        if self.quant_type != "nf4":
            self._check_quant_type()
            return self._forward_quantized(x, self.compute_dtype)
        if self.qweight is None:
            if not self.weight.is_cuda:
                raise RuntimeError("quanta_b200 layers run on CUDA only: move the module to a GPU first")
            self.quantize_()
        xin = x if x.dtype == self.compute_dtype else x.to(self.compute_dtype)
        y = linear_nf4a16(xin, self.qweight, self.scale, self.bias, blocksize=self.blocksize,
                          out_features=self.out_features, in_features=self.in_features)
        return y if y.dtype == x.dtype or not x.is_floating_point() else y.to(x.dtype)
