#!/usr/bin/env python
"""Debug: role timeline of the fused skim+decode kernel (needs a -DFRB_DEC_TIMING build: tools/build_variants.py
timing=-DFRB_DEC_TIMING, run with FRB_LIB_PATH=flac_raster_b200/lib/var_timing.so).  argv: [audio] [idx] [n_tiles]"""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from flac_raster_b200 import _native as nat, synth
from flac_raster_b200.engine import Engine, tile_grid
dev = torch.device("cuda", 0)
eng = Engine(dev)
L = nat.lib()
raster = synth.sentinel2_like(10980, 10980, 8, device=dev)
tiles = tile_grid(10980, 10980, 1024)
ntl = [int(a) for a in sys.argv[1:] if a.isdigit()]
if ntl:
    tiles = tiles[:ntl[0]]
enc = eng.encode_tiles(raster, tiles, 5)
payload = torch.cat([enc.payload, torch.zeros(64, dtype=torch.uint8, device=dev)])
dbg = (C.c_ulonglong * 16)()
out = torch.zeros_like(raster)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(3):
    ev0.record()
    L.frb_debug_decode_timing(None, 1)
    if len(sys.argv) > 1 and sys.argv[1] == "audio":
        audio, base, st = eng.decode_streams(payload, enc.offsets, enc.sizes, enc.n_samples, enc.sample_rates, 8, 16, 4096)
    else:
        st = eng.decode_tiles(payload, enc.offsets, enc.sizes, tiles, enc.sample_rates, enc.minmax, 32767.0, out, 16, 4096,
                              index=enc.index() if "idx" in sys.argv else None)
    ev1.record()
    torch.cuda.synchronize()
    L.frb_debug_decode_timing(dbg, 0)
    print("tiles", len(tiles), "step %.3f ms" % ev0.elapsed_time(ev1), end="  ")
    t0 = dbg[0]
    print("skim end %.3f ms, decode end %.3f ms, first decode CTA start %.3f; per-channel decode end:" % ((dbg[1]-t0)/1e6, (dbg[2]-t0)/1e6, (dbg[11]-t0)/1e6),
          " ".join("%.2f" % ((dbg[3+c]-t0)/1e6) for c in range(8)), "status", st[:6])
