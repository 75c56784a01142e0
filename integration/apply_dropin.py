#!/usr/bin/env python
"""Apply the drop-in described in INTEGRATION.md to a copy of the reference package.

    python integration/apply_dropin.py <reference Quanta dir> <output dir>

creates ``<output dir>/Quanta`` = the reference package plus

  1. ``Quanta/backends/cuda/{__init__,quantization}.py``: the CUDA entry the
     reference's dispatcher looks for (Quanta/backends/__init__.py:16-26);
  2. a device check at the top of the four public functions of ``Quanta/functional/quantization.py``
     (:7, :20, :33, :53) and of ``pack_4bit_tensor`` / ``unpack_4bit_tensor`` (``Quanta/utils/utils.py:23,37``
     and its duplicate ``utils/tensor_utils.py:6,20``) that routes CUDA tensors to quanta_b200
      — nothing in the reference routes these through ``Quanta.backends``, and its own
     CUDA tests (Quanta/tests/test_quantization.py:34-124) call them directly.

Nothing else is touched; CPU tensors take the reference's own code, byte for byte.  Used by
tests/test_reference_dropin.py; the input is never modified.
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

FUNCTIONAL_PATCHES = {
    # function name -> (first tensor argument, forwarded call)
    "quantize_4bit": ("tensor", "_q.quantize_4bit(tensor, quant_type, per_channel)"),
    "quantize_8bit": ("tensor", "_q.quantize_8bit(tensor, quant_type, per_channel)"),
    "dequantize_8bit": ("q_tensor", "_q.dequantize_8bit(q_tensor, scale_or_levels, zero_point_or_bias, quant_type)"),
    "dequantize_4bit": ("q_tensor", "_q.dequantize_4bit(q_tensor, scale_or_levels, zero_point_or_bias, quant_type)"),
}
UTILS_PATCHES = {
    "pack_4bit_tensor": ("tensor", "_u.pack_4bit_tensor(tensor)"),
    "unpack_4bit_tensor": ("packed_tensor", "_u.unpack_4bit_tensor(packed_tensor)"),
}


def _patch_functions(path, patches, import_line):
    src = open(path).read()
    for name, (arg, call) in patches.items():
        # def name(...):\n    """docstring"""\n  -> insert the device check after the docstring
        pat = re.compile(r"(def %s\([^)]*\):\n(?:    \"\"\".*?\"\"\"\n)?)" % re.escape(name), re.S)
        m = pat.search(src)
        if not m:
            raise RuntimeError(f"{path}: def {name} not found")
        check = (f"    if getattr({arg}, 'is_cuda', False):            # quanta_b200 drop-in (INTEGRATION.md §3)\n"
                 f"        {import_line}\n"
                 f"        return {call}\n")
        src = src[:m.end()] + check + src[m.end():]
    open(path, "w").write(src)


def apply(ref_pkg, out_dir):
    dst = os.path.join(out_dir, "Quanta")
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(ref_pkg, dst, ignore=shutil.ignore_patterns("__pycache__"))
    cuda_dir = os.path.join(dst, "backends", "cuda")
    os.makedirs(cuda_dir, exist_ok=True)
    for f in ("__init__.py", "quantization.py"):
        shutil.copy(os.path.join(HERE, "Quanta", "backends", "cuda", f), os.path.join(cuda_dir, f))
    _patch_functions(os.path.join(dst, "functional", "quantization.py"), FUNCTIONAL_PATCHES,
                     "from quanta_b200.functional import quantization as _q")
    for f in ("utils.py", "tensor_utils.py"):
        p = os.path.join(dst, "utils", f)
        if os.path.exists(p):
            _patch_functions(p, UTILS_PATCHES, "from quanta_b200.utils import utils as _u")
    return dst


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    print(apply(sys.argv[1], sys.argv[2]))
