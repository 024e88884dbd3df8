"""Encode step and kernel times of C3 and C5 with the library named by FRB_LIB_PATH (kernel variant A/B runs, tools/build_variants.py):
median over steps of the step's device time and of the library's bracketed kernels.  One line per workload."""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from flac_raster_b200 import synth
from flac_raster_b200.engine import Engine, tile_grid

eng = Engine(0)
L = eng.L
dev = torch.device("cuda", 0)
tag = os.path.basename(os.environ.get("FRB_LIB_PATH", "default"))
for wl in (sys.argv[1:] or ["c3", "c5"]):
    nb, H, W = bench.scene_shape(wl, 1)
    ts = bench.WORKLOADS[wl][2]
    tiles = tile_grid(H, W, ts)
    r = bench.make_rows(wl, dev, 0, H, 1)
    L.frb_profile_enable(1)
    acc, steps = {}, []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(8):
        torch.cuda.synchronize()
        ev0.record()
        enc = eng.encode_tiles(r, tiles, bench.WORKLOADS[wl][1])
        ev1.record()
        torch.cuda.synchronize()
        if it >= 3:
            steps.append(ev0.elapsed_time(ev1))
            for which, name in ((4, "stats"), (0, "code"), (2, "emit")):
                ms = C.c_float(0)
                if L.frb_profile_last_ms(which, C.byref(ms)) == 0:
                    acc.setdefault(name, []).append(ms.value)
    print(tag, wl, "step", round(float(np.median(steps)), 3), {k: round(float(np.median(v)), 3) for k, v in acc.items()}, "bytes", int(enc.payload.numel()), flush=True)
    del r, enc
    torch.cuda.empty_cache()
