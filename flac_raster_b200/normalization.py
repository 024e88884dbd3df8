"""Drop-in for flac_raster.normalization (reference src/flac_raster/normalization.py).

Same names, signatures and results; the elementwise work and the min/max
reductions run as sm_100a CUDA kernels (csrc/frb_normalize.cuh) through the C
ABI.  numpy arrays in -> numpy arrays out (H2D/D2H inside the call); torch CUDA
tensors in -> torch CUDA tensors out (no copies).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

logger = logging.getLogger("flac_raster.normalization")

_SUPPORTED = ("uint8", "int8", "uint16", "int16", "uint32", "int32", "float32", "float64")


@dataclass
class NormalizationParams:
    """Parameters needed for reversible normalization (normalization.py:27-56)."""

    data_min: float
    data_max: float
    original_dtype: str
    bits_per_sample: int
    scale_factor: int

    def to_dict(self) -> dict:
        return {
            "data_min": self.data_min,
            "data_max": self.data_max,
            "original_dtype": self.original_dtype,
            "bits_per_sample": self.bits_per_sample,
            "scale_factor": self.scale_factor,
        }

    @classmethod
    def from_dict(cls, d: dict) -> "NormalizationParams":
        return cls(
            data_min=d["data_min"],
            data_max=d["data_max"],
            original_dtype=d["original_dtype"],
            bits_per_sample=d["bits_per_sample"],
            scale_factor=d.get("scale_factor", 32767),
        )


def get_dtype_info(dtype) -> Tuple[Optional[float], Optional[float], bool]:
    """(min, max, is_integer) of a dtype (normalization.py:59-75)."""
    dtype = np.dtype(dtype)
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        return float(info.min), float(info.max), True
    if np.issubdtype(dtype, np.floating):
        return None, None, False
    raise ValueError(f"Unsupported dtype: {dtype}")


def audio_params_for(shape, dtype) -> Tuple[int, int]:
    """(sample_rate, bits_per_sample) from a shape and dtype (normalization.py:78-123)."""
    dtype = np.dtype(dtype)
    if dtype in (np.uint8, np.int8, np.uint16, np.int16):
        bits = 16
    elif dtype in (np.uint32, np.int32, np.float32, np.float64):
        bits = 24
    else:
        logger.warning(f"Unknown dtype {dtype}, defaulting to 24-bit")
        bits = 24
    shape = tuple(shape)
    if len(shape) >= 2:
        total = int(shape[-2]) * int(shape[-1])
    else:
        total = int(np.prod(shape)) if len(shape) else 1
    if total < 1_000_000:
        rate = 44100
    elif total < 10_000_000:
        rate = 48000
    elif total < 100_000_000:
        rate = 96000
    else:
        rate = 192000
    return rate, bits


def sample_rates_for_pixel_counts(npx: np.ndarray) -> np.ndarray:
    """Vectorised sample-rate rule of audio_params_for (normalization.py:113-120) for an array of pixel counts."""
    npx = np.asarray(npx, dtype=np.int64)
    return np.select([npx < 1_000_000, npx < 10_000_000, npx < 100_000_000], [44100, 48000, 96000], 192000).astype(np.uint32)


def calculate_audio_params(data, dtype) -> Tuple[int, int]:
    """Same as the reference: sample rate by pixel count, bit depth by dtype."""
    return audio_params_for(tuple(data.shape), dtype)


def _scale_for(bits_per_sample: int) -> int:
    return 32767 if bits_per_sample == 16 else 8388607 if bits_per_sample == 24 else 2147483647


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def normalize_to_audio(data, bits_per_sample: int, data_min: float = None, data_max: float = None):
    """normalize_to_audio (normalization.py:126-202) on the GPU.

    data: numpy array or torch CUDA tensor of any supported dtype and shape.
    Returns (audio, NormalizationParams); audio is int16 for 16-bit, int32
    otherwise, same shape and container type as the input.
    """
    import torch
    from . import _native as nat

    nat.require_cuda()
    L = nat.lib()
    is_t = _is_torch(data)
    if is_t:
        dt_name = str(data.dtype).replace("torch.", "")
        dev = data if data.is_cuda else data.cuda()
        dev = dev.contiguous()
        shape = tuple(data.shape)
    else:
        data = np.asarray(data)
        dt_name = str(data.dtype)
        shape = data.shape
    if dt_name not in _SUPPORTED:
        raise ValueError(f"Unsupported dtype: {dt_name}")
    if not is_t:
        host = np.ascontiguousarray(data)
        dev = torch.from_numpy(host.reshape(-1).view(np.uint8)).cuda()
    n = int(np.prod(shape)) if len(shape) else 1
    code = nat.DTYPE_CODES[dt_name]
    stream = torch.cuda.current_stream().cuda_stream
    if data_min is None or data_max is None:
        if n == 0:
            raise ValueError("zero-size array to reduction operation fmin which has no identity")
        mm = torch.empty(2, dtype=torch.float64, device=dev.device)
        nat.check(L.frb_minmax_flat(dev.data_ptr(), code, n, mm.data_ptr(), stream), "frb_minmax_flat")
        mmh = mm.cpu().numpy()
        if data_min is None:
            data_min = float(mmh[0])
        if data_max is None:
            data_max = float(mmh[1])
    if data_max <= data_min:
        logger.warning(f"Data has no range (min={data_min}, max={data_max}), using zeros")
    scale = _scale_for(bits_per_sample)
    out16 = bits_per_sample == 16
    out = torch.empty(n, dtype=torch.int16 if out16 else torch.int32, device=dev.device)
    nat.check(L.frb_normalize_flat(dev.data_ptr(), code, n, float(data_min), float(data_max), int(bits_per_sample),
                                   out.data_ptr(), int(out16), stream), "frb_normalize_flat")
    params = NormalizationParams(data_min=data_min, data_max=data_max, original_dtype=dt_name,
                                 bits_per_sample=bits_per_sample, scale_factor=scale)
    if is_t:
        return out.reshape(shape), params
    return out.cpu().numpy().reshape(shape), params


def denormalize_from_audio(audio_data, params: NormalizationParams):
    """denormalize_from_audio (normalization.py:205-253) on the GPU."""
    import torch
    from . import _native as nat

    nat.require_cuda()
    L = nat.lib()
    is_t = _is_torch(audio_data)
    if is_t:
        a_name = str(audio_data.dtype).replace("torch.", "")
        dev = (audio_data if audio_data.is_cuda else audio_data.cuda()).contiguous()
        shape = tuple(audio_data.shape)
    else:
        audio_data = np.asarray(audio_data)
        a_name = str(audio_data.dtype)
        shape = audio_data.shape
    # scale by audio dtype (normalization.py:222-232)
    if a_name == "int16":
        kind, scale = 0, 32767.0
    elif a_name == "int32":
        kind, scale = 1, float(params.scale_factor)
    elif a_name in ("float32", "float64"):
        kind, scale = 2, 1.0
        if a_name == "float32":
            if is_t:
                dev = dev.double()
            else:
                audio_data = audio_data.astype(np.float64)
    else:
        raise ValueError(f"Unsupported audio dtype: {a_name}")
    if not is_t:
        host = np.ascontiguousarray(audio_data)
        dev = torch.from_numpy(host.reshape(-1).view(np.uint8)).cuda()
    out_name = str(np.dtype(params.original_dtype))
    if out_name not in _SUPPORTED:
        raise ValueError(f"Unsupported dtype: {out_name}")
    n = int(np.prod(shape)) if len(shape) else 1
    itemsize = np.dtype(out_name).itemsize
    out = torch.empty(max(n * itemsize, 1), dtype=torch.uint8, device=dev.device)
    stream = torch.cuda.current_stream().cuda_stream
    nat.check(L.frb_denormalize_flat(dev.data_ptr(), kind, n, float(params.data_min), float(params.data_max), scale,
                                     out.data_ptr(), nat.DTYPE_CODES[out_name], stream), "frb_denormalize_flat")
    if is_t:
        from .engine import TORCH_DTYPES
        return out[:n * itemsize].view(TORCH_DTYPES[out_name]).reshape(shape)
    return out[:n * itemsize].cpu().numpy().view(out_name).reshape(shape)


def estimate_precision_loss(original_dtype, data_min: float, data_max: float, bits_per_sample: int) -> dict:
    """Pure arithmetic, as the reference (normalization.py:256-303)."""
    dtype = np.dtype(original_dtype)
    data_range = data_max - data_min
    if bits_per_sample == 16:
        levels = 65534
    elif bits_per_sample == 24:
        levels = 16777214
    else:
        levels = 4294967294
    max_error = data_range / levels
    rel = (max_error / data_range) * 100 if data_range > 0 else 0.0
    is_lossless = False
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        is_lossless = (int(info.max) - int(info.min)) <= levels
    return {
        "max_absolute_error": max_error,
        "relative_error_percent": rel,
        "quantization_levels": levels,
        "is_lossless": is_lossless,
        "bits_per_sample": bits_per_sample,
    }
