"""Drop-in CUDA implementation of the 4-bit pack helpers of ``Quanta.utils``
(Quanta/utils/utils.py:23-54; code-identical copy in utils/tensor_utils.py:6-37)."""
from __future__ import annotations

import torch

from .. import _host, _lib


def pack_4bit_tensor(tensor):
    """Pack 4-bit values into a tensor with half the size.  Returns
    ``(packed flat uint8 of ceil(numel/2), origin_shape)`` like utils.py:23-35:
    even index -> low nibble, one zero pad if numel is odd, inputs > 15 are not
    masked."""
    if tensor.dtype != torch.uint8:
        raise ValueError("Input tensor must be uint8")
    _host.require_cuda(tensor)
    origin_shape = tensor.shape
    q = tensor.detach().reshape(-1)
    if not q.is_contiguous():
        q = q.contiguous()
    n = q.numel()
    packed = torch.empty((n + 1) // 2, dtype=torch.uint8, device=q.device)
    if n:
        with _host.device_guard(q.device):
            st = _lib.lib().quanta_pack4(q.data_ptr(), n, packed.data_ptr(), _host.stream_ptr(q.device))
        _lib.check(st, "quanta_pack4")
    return packed, origin_shape


def unpack_4bit_tensor(packed_tensor):
    """Unpack a tensor where each byte contains two 4-bit values; returns the
    flat uint8 tensor of 2*len (pad element included; utils.py:37-48)."""
    _host.require_cuda(packed_tensor, "packed_tensor")
    p = packed_tensor.detach().reshape(-1)
    if p.dtype != torch.uint8:
        p = p.to(torch.uint8)
    if not p.is_contiguous():
        p = p.contiguous()
    out = torch.empty(p.numel() * 2, dtype=torch.uint8, device=p.device)
    if p.numel():
        with _host.device_guard(p.device):
            st = _lib.lib().quanta_unpack4(p.data_ptr(), p.numel(), out.data_ptr(), _host.stream_ptr(p.device))
        _lib.check(st, "quanta_unpack4")
    return out


def tensor_bits_to_bytes(tensor, bits):
    """Convert tensor size from bits to bytes (utils.py:50-54)."""
    n = 1
    for d in tensor.shape:
        n *= int(d)
    return (n * bits + 7) // 8


# --------------------------------------------------------------------------
# Row N3: the formats either side of the path — .qtn / .pt files and precision conversion
# (Quanta/utils/utils.py:60-208, :216-330).  File I/O is host work, like in the reference.
# --------------------------------------------------------------------------

def _np(v):
    import numpy as np
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu()
        return (v.float() if v.dtype in (torch.bfloat16, torch.float16) else v).numpy()
    return np.asarray(v)


def _with_suffix(path, suffix):
    return path if path.endswith(suffix) else path + suffix


def _count(shape):
    n = 1
    for d in shape:
        n *= int(d)
    return n


def save_quantized_tensor(q_tensor, scale, zero_point, params, file_path):
    """Write a ``.qtn`` file, byte-compatible with the reference writer (utils.py:60-108):
    ``u64le(len(header)) | header JSON {bits, scheme, type, shape, dtype} | codes | scale | zero point`` (raw bytes).
    Non-scalar scales (per_channel / blockwise) add ``scale_shape`` / ``zero_point_shape`` (and ``blocksize``) to the
    header so that :func:`load_quantized_tensor` can restore them — the reference's loader reads one float each."""
    import json
    import os
    import struct
    file_path = _with_suffix(file_path, ".qtn")
    os.makedirs(os.path.dirname(os.path.abspath(file_path)), exist_ok=True)
    import numpy as np
    # the loader (like the reference's, utils.py:159-163) reads codes as uint8 and the parameters as float32: write
    # exactly that, whatever the caller holds (a Python float from QuantizationState.load_state, a float16 /
    # float64 tensor ...).  The reference writes the native dtype and then cannot read the file back.
    payload = [np.ascontiguousarray(_np(q_tensor), dtype=np.uint8),
               np.ascontiguousarray(_np(scale), dtype=np.float32),
               np.ascontiguousarray(_np(zero_point), dtype=np.float32)]
    header = {"bits": params.get("bits", 8), "scheme": params.get("scheme", "symmetric"),
              "type": params.get("type", "linear"), "shape": list(q_tensor.shape),
              "dtype": str(q_tensor.dtype).replace("torch.", "", 1)}
    if payload[1].size > 1 or payload[2].size > 1:
        header.update(scale_shape=list(payload[1].shape), zero_point_shape=list(payload[2].shape))
        if params.get("blocksize") is not None:
            header["blocksize"] = int(params["blocksize"])
    text = json.dumps(header)
    with open(file_path, "wb") as out:
        out.write(struct.pack("<Q", len(text)) + text.encode("utf-8"))
        for part in payload:
            out.write(part.tobytes())
    return file_path


def load_quantized_tensor(file_path, device=None):
    """Read a ``.qtn`` file (utils.py:110-165) -> ``(q_tensor, scale, zero_point, metadata)``.  Files written by the
    reference (scalar float32 scale / zero point) load exactly as the reference loads them; files carrying
    ``scale_shape`` restore the parameter arrays.  ``device`` moves the result (e.g. ``"cuda"``)."""
    import json
    import struct
    import numpy as np

    def take(src, dtype, shape):
        n = _count(shape)
        return np.frombuffer(src.read(n * np.dtype(dtype).itemsize), dtype=dtype).copy().reshape(shape)

    with open(_with_suffix(file_path, ".qtn"), "rb") as src:
        (header_len,) = struct.unpack("<Q", src.read(8))
        metadata = json.loads(src.read(header_len).decode("utf-8"))
        q = torch.from_numpy(take(src, np.uint8, tuple(metadata["shape"])))
        if "scale_shape" in metadata:
            scale = torch.from_numpy(take(src, np.float32, tuple(metadata["scale_shape"])))
            zero_point = torch.from_numpy(take(src, np.float32, tuple(metadata["zero_point_shape"])))
        else:
            scale = torch.tensor(take(src, np.float32, ()).item())
            zero_point = torch.tensor(take(src, np.float32, ()).item())
    if device is not None:
        q, scale, zero_point = (t.to(device) for t in (q, scale, zero_point))
    return q, scale, zero_point, metadata


def save_quantized_tensor_torch(q_tensor, scale, zero_point, params, file_path):
    """``torch.save`` of ``{q_tensor, scale, zero_point, params}`` (utils.py:166-186)."""
    file_path = _with_suffix(file_path, ".pt")
    torch.save(dict(q_tensor=q_tensor, scale=scale, zero_point=zero_point, params=params), file_path)
    return file_path


def load_quantized_tensor_torch(file_path, map_location=None):
    """Inverse of :func:`save_quantized_tensor_torch` (utils.py:188-208)."""
    blob = torch.load(_with_suffix(file_path, ".pt"), map_location=map_location, weights_only=False)
    return tuple(blob[k] for k in ("q_tensor", "scale", "zero_point", "params"))


def convert_precision(q_tensor, source_params, target_bits, target_type="linear", target_scheme=None):
    """Dequantize with the source parameters and quantize again at ``target_bits`` / ``target_type``
    (utils.py:216-279) — both steps on the GPU, nothing leaves the device.  Returns
    ``(q, scale_or_levels, zero_point_or_bias, new_params)``."""
    from ..functional import quantization as fq
    decode = {8: fq.dequantize_8bit, 4: fq.dequantize_4bit}
    encode = {8: fq.quantize_8bit, 4: fq.quantize_4bit}
    source_bits = source_params.get("bits", 8)
    if source_bits not in decode:
        raise ValueError(f"Unsupported source bit depth: {source_bits}")
    if target_bits not in encode:
        raise ValueError(f"Unsupported target bit depth: {target_bits}")
    fused = _convert_linear_fused(q_tensor, source_params, target_bits, target_type)
    if fused is not None:
        q, first, second = fused
    else:
        full = decode[source_bits](q_tensor, source_params.get("scale"), source_params.get("zero_point"),
                                   quant_type=source_params.get("type", "linear"))
        q, first, second = encode[target_bits](full, quant_type=target_type)
    scheme = target_scheme if target_scheme is not None else source_params.get("scheme", "symmetric")
    return q, first, second, {"bits": target_bits, "type": target_type, "scheme": scheme, "scale": first,
                              "zero_point": second, "shape": tuple(q.shape)}


def _convert_linear_fused(q_tensor, source_params, target_bits, target_type):
    """linear -> linear with ONE (scale, zero_point) for the tensor: the new code is a function of the old code, so
    the conversion is a 256-entry table applied to the codes (``quanta_convert_linear``: 3 B/element instead of a
    dequantize + quantize through fp32).  Returns None when the case does not apply (other types, per-channel /
    blockwise parameters)."""
    from .. import _host, _lib
    if source_params.get("type", "linear") != "linear" or target_type != "linear":
        return None
    if not (isinstance(q_tensor, torch.Tensor) and q_tensor.is_cuda and q_tensor.dtype == torch.uint8 and q_tensor.numel() > 0):
        return None
    dev = q_tensor.device
    params = []
    for key in ("scale", "zero_point"):
        v = source_params.get(key)
        if v is None:
            return None
        v = torch.as_tensor(v, dtype=torch.float32, device=dev)
        if v.numel() != 1:
            return None
        params.append(v.reshape(1).contiguous())
    q = q_tensor.detach()
    if not q.is_contiguous():
        q = q.contiguous()
    with _host.device_guard(dev):
        out = torch.empty(q.shape, dtype=torch.uint8, device=dev)
        scale, zp = torch.empty((), dtype=torch.float32, device=dev), torch.empty((), dtype=torch.float32, device=dev)
        ws = _host.workspace(dev, 256)
        st = _lib.lib().quanta_convert_linear(q.data_ptr(), q.numel(), params[0].data_ptr(), params[1].data_ptr(), int(target_bits),
                                              out.data_ptr(), scale.data_ptr(), zp.data_ptr(), ws.data_ptr(), ws.numel(),
                                              _host.stream_ptr(dev))
    _lib.check(st, "quanta_convert_linear")
    return out, scale, zp


def convert_8bit_to_4bit(q_tensor, source_params, target_type="linear"):
    """utils.py:281-293."""
    return convert_precision(q_tensor, source_params, 4, target_type)


def convert_4bit_to_8bit(q_tensor, source_params, target_type="linear"):
    """utils.py:295-307."""
    return convert_precision(q_tensor, source_params, 8, target_type)


_HARDWARE = {"cpu": (8, "linear"), "gpu": (8, "linear"), "mobile": (4, "nf4"), "edge": (4, "linear")}


def optimize_for_target_hardware(q_tensor, source_params, target_hardware):
    """The reference's hardware table (utils.py:309-337): cpu / gpu -> 8-bit linear, mobile -> nf4, edge -> 4-bit
    linear, anything else -> 8-bit linear."""
    bits, kind = _HARDWARE.get(target_hardware, (8, "linear"))
    return convert_precision(q_tensor, source_params, bits, kind)
