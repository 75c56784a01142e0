"""Mirror of ``Quanta.functional`` (Quanta/functional/__init__.py:5-16 exports
only the quantization primitives), plus the batched blockwise entry points.
``Quanta.functional.base.BaseQuantizer`` (convention C) lives in ``.base``, like in the reference."""
from .quantization import (quantize_8bit, quantize_4bit, dequantize_8bit, dequantize_4bit,
                           quantize_4bit_many, quantize_8bit_many, quantize_nf4_many,
                           dequantize_4bit_many, dequantize_8bit_many)

__all__ = ["quantize_8bit", "quantize_4bit", "dequantize_8bit", "dequantize_4bit",
           "quantize_4bit_many", "quantize_8bit_many", "quantize_nf4_many",
           "dequantize_4bit_many", "dequantize_8bit_many"]
