#!/usr/bin/env python
"""Does the HBM-bound normalise co-run with the issue-bound analysis when its grid is capped (FRB_MAP_CTAS)?"""
import sys, time, os
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from flac_raster_b200 import synth
from flac_raster_b200.engine import Engine, tile_grid
from flac_raster_b200.normalization import audio_params_for
dev = torch.device("cuda", 0)
raster = synth.sentinel2_like(10980, 10980, 8, device=dev)
tiles = tile_grid(10980, 10980, 1024)
e1, e2 = Engine(dev), Engine(dev)
rates = np.array([audio_params_for((int(t["h"]), int(t["w"])), "uint16")[0] for t in tiles], dtype=np.uint32)
def T(fn, n=4):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
audio, base, npx, mm, bits = e1.normalize_tiles(raster, tiles)
t_norm = T(lambda: e2.normalize_tiles(raster, tiles))
t_an = T(lambda: e1.encode_audio(audio, npx, base, rates, 8, 16, 5))
for prio in (0, -1):
    s1 = torch.cuda.Stream(priority=prio)
    def both():
        s1.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            e2.normalize_tiles(raster, tiles)
        e1.encode_audio(audio, npx, base, rates, 8, 16, 5)
        torch.cuda.current_stream().wait_stream(s1)
    t_both = T(both)
    print(f"FRB_MAP_CTAS={os.environ.get('FRB_MAP_CTAS','-')} prio={prio}: normalise {t_norm:.2f} ms, analyse+emit {t_an:.2f} ms, concurrently {t_both:.2f} ms (sum {t_norm+t_an:.2f})")
