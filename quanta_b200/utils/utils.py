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
        with torch.cuda.device(q.device):
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
        with torch.cuda.device(p.device):
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
        return v.detach().cpu().numpy()
    return np.asarray(v)


def save_quantized_tensor(q_tensor, scale, zero_point, params, file_path):
    """Write a ``.qtn`` file, byte-compatible with the reference writer (utils.py:60-108):
    8-byte little-endian header length, JSON metadata ``{bits, scheme, type, shape, dtype}``, then the raw
    bytes of the codes, the scale and the zero point.  Non-scalar scales (per_channel / blockwise) add
    ``scale_shape`` / ``zero_point_shape`` to the metadata so that :func:`load_quantized_tensor` can
    restore them (the reference's loader only reads one float each)."""
    import json
    import os
    os.makedirs(os.path.dirname(os.path.abspath(file_path)), exist_ok=True)
    if not file_path.endswith(".qtn"):
        file_path = file_path + ".qtn"
    q_np, s_np, z_np = _np(q_tensor), _np(scale), _np(zero_point)
    dtype_str = str(q_tensor.dtype)
    if dtype_str.startswith("torch."):
        dtype_str = dtype_str[6:]
    metadata = {"bits": params.get("bits", 8), "scheme": params.get("scheme", "symmetric"),
                "type": params.get("type", "linear"), "shape": list(q_tensor.shape), "dtype": dtype_str}
    if s_np.size > 1 or z_np.size > 1:
        metadata["scale_shape"] = list(s_np.shape)
        metadata["zero_point_shape"] = list(z_np.shape)
        if params.get("blocksize") is not None:
            metadata["blocksize"] = int(params["blocksize"])
    with open(file_path, "wb") as f:
        header = json.dumps(metadata)
        f.write(len(header).to_bytes(8, byteorder="little"))
        f.write(header.encode("utf-8"))
        f.write(q_np.tobytes())
        f.write(s_np.tobytes())
        f.write(z_np.tobytes())
    return file_path


def load_quantized_tensor(file_path, device=None):
    """Read a ``.qtn`` file (utils.py:110-165): returns ``(q_tensor, scale, zero_point, metadata)``; files
    written by the reference (scalar float32 scale / zero point) load exactly as the reference loads them,
    files with ``scale_shape`` restore the parameter arrays.  ``device`` moves the result (e.g. ``"cuda"``)."""
    import json
    import numpy as np
    if not file_path.endswith(".qtn"):
        file_path = file_path + ".qtn"
    with open(file_path, "rb") as f:
        header_len = int.from_bytes(f.read(8), byteorder="little")
        metadata = json.loads(f.read(header_len).decode("utf-8"))
        shape = tuple(metadata["shape"])
        count = int(np.prod(shape)) if len(shape) else 1
        q = torch.from_numpy(np.frombuffer(f.read(count), dtype=np.uint8).copy().reshape(shape))
        if "scale_shape" in metadata:
            s_shape, z_shape = tuple(metadata["scale_shape"]), tuple(metadata["zero_point_shape"])
            ns, nz = int(np.prod(s_shape)) if len(s_shape) else 1, int(np.prod(z_shape)) if len(z_shape) else 1
            scale = torch.from_numpy(np.frombuffer(f.read(4 * ns), dtype=np.float32).copy().reshape(s_shape))
            zero_point = torch.from_numpy(np.frombuffer(f.read(4 * nz), dtype=np.float32).copy().reshape(z_shape))
        else:
            scale = torch.tensor(np.frombuffer(f.read(4), dtype=np.float32).copy().item())
            zero_point = torch.tensor(np.frombuffer(f.read(4), dtype=np.float32).copy().item())
    if device is not None:
        q, scale, zero_point = q.to(device), scale.to(device), zero_point.to(device)
    return q, scale, zero_point, metadata


def save_quantized_tensor_torch(q_tensor, scale, zero_point, params, file_path):
    """``torch.save`` of ``{q_tensor, scale, zero_point, params}`` (utils.py:166-186)."""
    if not file_path.endswith(".pt"):
        file_path = file_path + ".pt"
    torch.save({"q_tensor": q_tensor, "scale": scale, "zero_point": zero_point, "params": params}, file_path)
    return file_path


def load_quantized_tensor_torch(file_path, map_location=None):
    """Inverse of :func:`save_quantized_tensor_torch` (utils.py:188-208)."""
    if not file_path.endswith(".pt"):
        file_path = file_path + ".pt"
    d = torch.load(file_path, map_location=map_location, weights_only=False)
    return d["q_tensor"], d["scale"], d["zero_point"], d["params"]


def convert_precision(q_tensor, source_params, target_bits, target_type="linear", target_scheme=None):
    """Dequantize with the source parameters and quantize again at ``target_bits`` / ``target_type``
    (utils.py:216-279) — both steps on the GPU, nothing leaves the device.  Returns
    ``(q, scale_or_levels, zero_point_or_bias, new_params)``."""
    from ..functional.quantization import quantize_8bit, quantize_4bit, dequantize_8bit, dequantize_4bit
    source_bits = source_params.get("bits", 8)
    source_type = source_params.get("type", "linear")
    if target_scheme is None:
        target_scheme = source_params.get("scheme", "symmetric")
    if source_bits == 8:
        fp = dequantize_8bit(q_tensor, source_params.get("scale"), source_params.get("zero_point"), quant_type=source_type)
    elif source_bits == 4:
        fp = dequantize_4bit(q_tensor, source_params.get("scale"), source_params.get("zero_point"), quant_type=source_type)
    else:
        raise ValueError(f"Unsupported source bit depth: {source_bits}")
    if target_bits == 8:
        new_q, new_scale, new_zp = quantize_8bit(fp, quant_type=target_type)
    elif target_bits == 4:
        new_q, new_scale, new_zp = quantize_4bit(fp, quant_type=target_type)
    else:
        raise ValueError(f"Unsupported target bit depth: {target_bits}")
    new_params = {"bits": target_bits, "type": target_type, "scheme": target_scheme, "scale": new_scale,
                  "zero_point": new_zp, "shape": tuple(new_q.shape)}
    return new_q, new_scale, new_zp, new_params


def convert_8bit_to_4bit(q_tensor, source_params, target_type="linear"):
    """utils.py:281-293."""
    return convert_precision(q_tensor, source_params, 4, target_type)


def convert_4bit_to_8bit(q_tensor, source_params, target_type="linear"):
    """utils.py:295-307."""
    return convert_precision(q_tensor, source_params, 8, target_type)


def optimize_for_target_hardware(q_tensor, source_params, target_hardware):
    """The reference's hardware table (utils.py:309-337): cpu / gpu -> 8-bit linear, mobile -> nf4, edge -> 4-bit linear."""
    hw = {"cpu": {"bits": 8, "type": "linear"}, "gpu": {"bits": 8, "type": "linear"},
          "mobile": {"bits": 4, "type": "nf4"}, "edge": {"bits": 4, "type": "linear"}}
    cfg = hw.get(target_hardware, {"bits": 8, "type": "linear"})
    return convert_precision(q_tensor, source_params, cfg["bits"], cfg["type"])
