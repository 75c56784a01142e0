"""Quanta/backends/cuda/quantization.py — the file the reference's dispatcher try-imports
(Quanta/backends/__init__.py:16-26) and does not ship.

It exports exactly the four names the dispatcher forwards to
(``quantize_8bit_cuda``, ``dequantize_8bit_cuda``, ``quantize_4bit_cuda``,
``dequantize_4bit_cuda``), implemented by the sm_100a kernels of quanta_b200 behind the
C-ABI ``quanta_backend_quantize`` / ``quanta_backend_dequantize`` (include/quanta_b200.h).
There is no fallback: if quanta_b200 or its shared library is missing this import raises,
and the dispatcher's ``except ImportError`` then reports ``CUDA_AVAILABLE = False`` exactly
as it does today.
"""
from quanta_b200.backends.cuda.quantization import (  # noqa: F401
    quantize_8bit_cuda,
    dequantize_8bit_cuda,
    quantize_4bit_cuda,
    dequantize_4bit_cuda,
)
