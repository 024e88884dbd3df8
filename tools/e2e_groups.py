#!/usr/bin/env python
"""e2e encode step (pinned host raster in, host frames out) against the number of pipeline stages."""
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
total = raster.numel() * 2
for div in (6, 11, 12, 16, 24):
    gb = total // div
    for _ in range(2):
        eng.encode_tiles_host(host_raster, tiles, 5, host_out=host_out, group_bytes=gb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        e = eng.encode_tiles_host(host_raster, tiles, 5, host_out=host_out, group_bytes=gb)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print(f"group_bytes=total/{div}: {ms:.1f} ms  {raster.numel()/ms/1e6:.2f} GS/s")
