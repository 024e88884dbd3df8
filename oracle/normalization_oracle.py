"""numpy restatement of the reference's sample mapping -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/flac_raster/normalization.py line by line (the
reference itself is plain numpy, so this *is* the algorithm, restated so it can
travel to the GPU box where /root/reference does not exist).  Pinned against
outputs of the real reference module in tests/golden/normalization_vectors.npz
(generator: tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import
this module; the product package never does.
"""
from __future__ import annotations

import numpy as np


def calculate_audio_params(shape, dtype):
    """normalization.py:78-123 -> (sample_rate, bits_per_sample)."""
    dtype = np.dtype(dtype)
    if dtype in (np.uint8, np.int8, np.uint16, np.int16):      # :92-97
        bits = 16
    else:                                                        # :98-104
        bits = 24
    total = shape[-2] * shape[-1] if len(shape) >= 2 else int(np.prod(shape))   # :108-111
    if total < 1_000_000:                                        # :113-120
        rate = 44100
    elif total < 10_000_000:
        rate = 48000
    elif total < 100_000_000:
        rate = 96000
    else:
        rate = 192000
    return rate, bits


def normalize_to_audio(data, bits_per_sample, data_min=None, data_max=None):
    """normalization.py:126-202 -> (audio, dict(params))."""
    if data_min is None:
        data_min = float(np.nanmin(data))                        # :149-151
    if data_max is None:
        data_max = float(np.nanmax(data))                        # :152-153
    data_range = 1.0 if data_max <= data_min else data_max - data_min   # :156-160
    with np.errstate(all="ignore"):
        x = data.astype(np.float64)                              # :164
        n = 2.0 * (x - data_min) / data_range - 1.0              # :165
        n = np.clip(n, -1.0, 1.0)                                # :168
        n[np.isnan(n)] = 0.0                                     # :171-174
        if bits_per_sample == 16:                                # :177-179
            scale = 32767
            audio = (n * scale).astype(np.int16)
        elif bits_per_sample == 24:                              # :180-183
            scale = 8388607
            audio = (n * scale).astype(np.int32)
        else:                                                    # :184-187
            scale = 2147483647
            audio = (n * scale).astype(np.int32)
    return audio, dict(data_min=data_min, data_max=data_max, original_dtype=str(data.dtype),
                       bits_per_sample=bits_per_sample, scale_factor=scale)


def denormalize_from_audio(audio, data_min, data_max, original_dtype, scale_factor):
    """normalization.py:205-253."""
    if audio.dtype == np.int16:                                  # :222-223
        scale = 32767.0
    elif audio.dtype == np.int32:                                # :224-227
        scale = float(scale_factor)
    elif audio.dtype in (np.float32, np.float64):                # :228-230
        scale = 1.0
    else:
        scale = float(scale_factor)
    with np.errstate(all="ignore"):
        n = audio.astype(np.float64) / scale                     # :235
        rng = data_max - data_min                                # :239
        x = (n + 1.0) / 2.0 * rng + data_min                     # :240
        dt = np.dtype(original_dtype)
        if np.issubdtype(dt, np.integer):                        # :245-247
            return np.round(x).astype(dt)
        return x.astype(dt)                                      # :249
