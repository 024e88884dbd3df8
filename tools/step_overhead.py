"""Where the fixed cost of a small encode step goes (a rank's share of C3 at N = 8: 16 tiles): device time of the step
(CUDA events), host time spent enqueueing it (perf_counter around encode_tiles with the stream left to run), and a
cProfile of the host side.  Prints a short report."""
import cProfile, pstats, io, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flac_raster_b200.engine import Engine, tile_grid
from flac_raster_b200 import synth

n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 16
eng = Engine(0)
H = W = 10980
tiles_all = tile_grid(H, W, 1024)
tiles = tiles_all[:n_tiles].copy()
rows = int((tiles["row_off"] + tiles["h"]).max())
sys.path.insert(0, str(ROOT))
import bench
r = bench.make_rows("c3", torch.device("cuda", 0), 0, rows)
for _ in range(5):
    enc = eng.encode_tiles(r, tiles, 5)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dev_ms, host_ms = [], []
for _ in range(20):
    torch.cuda.synchronize()
    ev0.record()
    t0 = time.perf_counter()
    enc = eng.encode_tiles(r, tiles, 5)
    t1 = time.perf_counter()
    ev1.record()
    torch.cuda.synchronize()
    dev_ms.append(ev0.elapsed_time(ev1)); host_ms.append(1e3 * (t1 - t0))
print(f"tiles {n_tiles}: device step {np.median(dev_ms):.3f} ms, host call {np.median(host_ms):.3f} ms (the call ends with the emit's status read-back, so it includes the device time)")
L = eng.L
L.frb_profile_enable(1)
enc = eng.encode_tiles(r, tiles, 5)
torch.cuda.synchronize()
import ctypes as C
ks = {}
for which, name in ((4, "k_enc_stats"), (0, "k_enc_code"), (2, "k_emit_frames"), (6, "analysis_total")):
    ms = C.c_float(0)
    if L.frb_profile_last_ms(which, C.byref(ms)) == 0:
        ks[name] = round(ms.value, 3)
L.frb_profile_enable(0)
print("kernels:", ks)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    enc = eng.encode_tiles(r, tiles, 5)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
print("\n".join(s.getvalue().splitlines()[:30]))
