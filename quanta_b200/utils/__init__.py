from .utils import pack_4bit_tensor, unpack_4bit_tensor, tensor_bits_to_bytes

__all__ = ["pack_4bit_tensor", "unpack_4bit_tensor", "tensor_bits_to_bytes"]
