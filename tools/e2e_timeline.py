#!/usr/bin/env python
"""Device timeline of the end-to-end encode pipeline (H2D / encode / D2H per stage) to find bubbles."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from flac_raster_b200 import synth
from flac_raster_b200.engine import Engine, tile_grid
dev = torch.device("cuda", 0)
eng = Engine(dev)
raster = synth.sentinel2_like(10980, 10980, 8, device=dev)
tiles = tile_grid(10980, 10980, 1024)
host = torch.empty(raster.numel() * 2, dtype=torch.uint8).pin_memory()
host.copy_(raster.reshape(-1).view(torch.uint8))
host_raster = host.view(raster.dtype).reshape(raster.shape)
host_out = torch.empty(1_500_000_000, dtype=torch.uint8).pin_memory()
for _ in range(2):
    eng.encode_tiles_host(host_raster, tiles, 5, host_out=host_out)
tl = []
torch.cuda.synchronize(); t0 = time.perf_counter()
eng.encode_tiles_host(host_raster, tiles, 5, host_out=host_out, timeline=tl)
print(f"wall {1e3 * (time.perf_counter() - t0):.1f} ms")
rows = {}
for name, g, t in tl:
    rows.setdefault(g, {})[name] = t
print(" g   h2d_begin  h2d_end  enc_begin  enc_end  d2h_begin  d2h_end")
for g in sorted(rows):
    r = rows[g]
    print(f"{g:2d} " + " ".join(f"{r.get(k, float('nan')):9.2f}" for k in ("h2d_begin", "h2d_end", "enc_begin", "enc_end", "d2h_begin", "d2h_end")))
print("host-side enqueue times (ms):")
for g in sorted(rows):
    r = rows[g]
    print(f"{g:2d} " + " ".join(f"{r.get('host_' + k, float('nan')):9.2f}" for k in ("h2d_begin", "h2d_end", "enc_begin", "enc_end", "d2h_begin", "d2h_end")))
