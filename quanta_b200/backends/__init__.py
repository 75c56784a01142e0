"""Mirror of ``Quanta.backends`` (Quanta/backends/__init__.py:42-127) with ONE
backend: CUDA.  The reference silently falls back to CPU when its optional
``.cuda.quantization`` import fails (:16-26); here a CPU tensor or a missing
extension raises instead (north_star: no multi-backend dispatch, no CPU
fallback)."""
from .cuda.quantization import quantize_8bit_cuda, dequantize_8bit_cuda, quantize_4bit_cuda, dequantize_4bit_cuda

CUDA_AVAILABLE = True


def quantize_8bit(tensor, per_channel=False, symmetric=True):
    return quantize_8bit_cuda(tensor, per_channel, symmetric)


def dequantize_8bit(q_tensor, scale, zero_point):
    return dequantize_8bit_cuda(q_tensor, scale, zero_point)


def quantize_4bit(tensor, per_channel=False, symmetric=True):
    return quantize_4bit_cuda(tensor, per_channel, symmetric)


def dequantize_4bit(q_tensor, scale, zero_point):
    return dequantize_4bit_cuda(q_tensor, scale, zero_point)
