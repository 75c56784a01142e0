from .utils import (pack_4bit_tensor, unpack_4bit_tensor, tensor_bits_to_bytes, save_quantized_tensor,
                    load_quantized_tensor, save_quantized_tensor_torch, load_quantized_tensor_torch, convert_precision,
                    convert_8bit_to_4bit, convert_4bit_to_8bit, optimize_for_target_hardware)

__all__ = ["pack_4bit_tensor", "unpack_4bit_tensor", "tensor_bits_to_bytes", "save_quantized_tensor",
           "load_quantized_tensor", "save_quantized_tensor_torch", "load_quantized_tensor_torch", "convert_precision",
           "convert_8bit_to_4bit", "convert_4bit_to_8bit", "optimize_for_target_hardware"]
