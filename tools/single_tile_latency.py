"""Decode latency of ONE C3 tile (8 bands x 1024^2: 2048 subframes) and of a rank's 16-tile share, device resident, with and
without the seek index: the regime where the launch time is a thread's dependent chain, not the issue rate."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from flac_raster_b200.engine import Engine, tile_grid

eng = Engine(0)
tiles_all = tile_grid(10980, 10980, 1024)
for n_tiles in (1, 16):
    tiles = tiles_all[:n_tiles].copy()
    rows = int((tiles["row_off"] + tiles["h"]).max())
    r = bench.make_rows("c3", torch.device("cuda", 0), 0, rows)
    enc = eng.encode_tiles(r, tiles, 5)
    pay = torch.cat([enc.payload, torch.zeros(64, dtype=torch.uint8, device=r.device)])
    out = torch.zeros_like(r)
    for name, idx in (("indexed", enc.index()), ("scan", None)):
        ts = []
        for rep in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            st = eng.decode_tiles(pay, enc.offsets, enc.sizes, tiles, enc.sample_rates, enc.minmax, 32767.0, out, 16, 4096, index=idx)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        assert list(st[:3]) == [0, 0, 0], st
        for t in tiles:
            r0, c0, h, w = (int(t[k]) for k in ("row_off", "col_off", "h", "w"))
            assert torch.equal(out[:, r0:r0 + h, c0:c0 + w].view(torch.int16), r[:, r0:r0 + h, c0:c0 + w].view(torch.int16))
        print(f"{n_tiles:3d} tile(s) {name:8s}: {np.median(ts[2:]):.3f} ms per call (min {min(ts[2:]):.3f})")
