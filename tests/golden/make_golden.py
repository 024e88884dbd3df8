"""Generate tests/golden/normalization_vectors.npz from the REAL reference module.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference's normalization.py imports nothing but numpy, so it is loaded by
path and called directly; inputs are seeded.  The .npz travels to the GPU box.
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src/flac_raster/normalization.py")
spec = importlib.util.spec_from_file_location("ref_normalization", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

DTYPES = ["uint8", "int8", "uint16", "int16", "uint32", "int32", "float32", "float64"]
N = 6000


def make_input(dt, case, rng):
    dt = np.dtype(dt)
    if np.issubdtype(dt, np.integer):
        info = np.iinfo(dt)
        if case == "full":
            x = rng.integers(info.min, info.max, size=N, endpoint=True, dtype=np.int64).astype(dt)
            x[0], x[1] = info.min, info.max
        elif case == "narrow":
            lo = max(info.min, -100)
            x = rng.integers(lo, min(info.max, lo + 200), size=N, dtype=np.int64).astype(dt)
        else:  # constant
            x = np.full(N, 7, dtype=dt)
    else:
        if case == "full":
            x = (rng.standard_normal(N) * 1000.0 + 250.0).astype(dt)
        elif case == "narrow":
            x = (rng.random(N) * 1e-3 + 5.0).astype(dt)
        elif case == "nan":
            x = (rng.standard_normal(N) * 10.0).astype(dt)
            x[::17] = np.nan
        else:
            x = np.full(N, -2.5, dtype=dt)
    return x


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    for dt in DTYPES:
        cases = ["full", "narrow", "constant"] + (["nan"] if dt.startswith("float") else [])
        sr, bits = ref.calculate_audio_params(np.zeros((3, 4), dtype=dt), np.dtype(dt))
        for case in cases:
            x = make_input(dt, case, rng)
            audio, p = ref.normalize_to_audio(x.reshape(-1, 3), bits)
            back = ref.denormalize_from_audio(audio, p)
            key = f"{dt}__{case}"
            out[key + "__in"] = x
            out[key + "__audio"] = audio
            out[key + "__back"] = back
            out[key + "__params"] = np.array([p.data_min, p.data_max, p.bits_per_sample, p.scale_factor], dtype=np.float64)
    # sample-rate table (normalization.py:113-120)
    shapes = [(512, 512), (1000, 1000), (1024, 1024), (3162, 3163), (10000, 10000), (10980, 10980)]
    out["audio_params_shapes"] = np.array(shapes)
    out["audio_params_rates"] = np.array([ref.calculate_audio_params(np.zeros((1, 1)).reshape(1, 1) if False else np.broadcast_to(np.zeros(1, dtype=np.uint16), s), np.uint16)[0] for s in shapes])
    dst = Path(__file__).with_name("normalization_vectors.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes", len(out), "arrays")


if __name__ == "__main__":
    sys.exit(main())
