from .quantization import quantize_8bit_cuda, dequantize_8bit_cuda, quantize_4bit_cuda, dequantize_4bit_cuda

__all__ = ["quantize_8bit_cuda", "dequantize_8bit_cuda", "quantize_4bit_cuda", "dequantize_4bit_cuda"]
