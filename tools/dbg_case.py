#!/usr/bin/env python
"""Debug: first subframe whose coding differs between the GPU encoder and the oracle for case K of the seeded sweep."""
import sys, importlib.util
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from flac_raster_b200 import _native as nat, flacfmt
from oracle import flac_oracle as fo
spec = importlib.util.spec_from_file_location("tgp", str(ROOT / "tests/test_gpu_parity.py")); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
want_case = int(sys.argv[1]) if len(sys.argv) > 1 else -1
rng = np.random.default_rng(20261018)
bad = 0
for case in range(36):
    ch = int(rng.integers(1, 9)) if case % 3 else 2
    bps = 16 if rng.random() < 0.7 else 32
    n = int(rng.choice([17, 4095, 4096, 4097, 8192, 12289, 20000, 33000]))
    level = int(rng.integers(0, 9))
    x, kinds = m._random_signal(rng, n, ch, bps)
    if want_case >= 0 and case != want_case:
        continue
    payload, fs = nat.host_encode(x, bps, 48000, level)
    oenc, ofs, od = fo.encode(x, bps, 48000, level, mid_side=(bps == 16), want_descs=True)
    osz = int(ofs.sum())
    same = bytes(payload) == oenc[len(oenc) - osz:]
    print(f"case {case}: n={n} ch={ch} bps={bps} level={level} same={same} gpu={len(payload)} oracle={osz}")
    if same:
        continue
    bad += 1
    si = flacfmt.StreamInfo(4096, 4096, 0, 0, 48000, ch, bps, n)
    dec, info, gd = fo.decode(flacfmt.build_header(si) + bytes(payload), want_descs=True)
    print("  gpu stream decodes to input:", np.array_equal(dec, x))
    shown = 0
    for a, b in zip(gd, od):
        keys = ("type", "order", "wasted", "precision", "shift", "coefs", "method", "partition_order", "params", "nbits", "ch_assign")
        if any(a[k] != b[k] for k in keys):
            print("  frame", a["frame"], "channel", a["channel"])
            for k in keys:
                if a[k] != b[k]:
                    print(f"    {k}: gpu={a[k]} oracle={b[k]}")
            col = x[a["frame"] * 4096:(a["frame"] + 1) * 4096, a["channel"]]
            print("    signal: min", col.min(), "max", col.max(), "nunique", len(np.unique(col)), "first", col[:6].tolist())
            shown += 1
            if shown >= 4:
                break
print("mismatching cases:", bad)
