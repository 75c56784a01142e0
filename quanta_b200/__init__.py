"""quanta_b200 — B200-native (sm_100a) implementation of the weight-quantization
hot path of ved1beta/Quanta, behind the reference's own Python API.

Layout mirrors the reference package for the hot path only:
  quanta_b200.functional            <- Quanta.functional.quantization ("linear")
  quanta_b200.utils                 <- Quanta.utils pack/unpack helpers
  quanta_b200.backends.cuda         <- the Quanta/backends/cuda entry the reference looks for
  quanta_b200.nn                    <- Quanta.nn Linear8bitLt / Linear4bit (dequant-GEMM forward)
"""
__version__ = "0.1.0"

from . import functional  # noqa: F401
from .functional import (quantize_8bit, quantize_4bit, dequantize_8bit, dequantize_4bit,  # noqa: F401
                         quantize_4bit_many, quantize_8bit_many, quantize_nf4_many,
                         dequantize_4bit_many, dequantize_8bit_many)
from .utils import pack_4bit_tensor, unpack_4bit_tensor  # noqa: F401
