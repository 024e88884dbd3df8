import sys, time, cProfile, pstats, tempfile
from pathlib import Path
sys.path.insert(0, '.')
import numpy as np, torch
from flac_raster_b200 import SpatialFLACEncoder, SpatialFLACStreamer
from flac_raster_b200.tiffio import write_geotiff
G = Path("tests/golden")
d = Path(tempfile.mkdtemp())
# C2: the reference's RGB sample, one tile
SpatialFLACEncoder(tile_size=512).encode(G / "sample_rgb.tif", d / "rgb.flac", streaming=True)
st = SpatialFLACStreamer(d / "rgb.flac")
for _ in range(5): st.get_tile_by_id(0)
ts = []
for _ in range(30):
    t0 = time.perf_counter(); st.get_tile_by_id(0); ts.append(time.perf_counter() - t0)
print("C2 get_tile_by_id ms: median %.3f min %.3f" % (1e3 * np.median(ts), 1e3 * min(ts)))
pr = cProfile.Profile(); pr.enable()
for _ in range(50): st.get_tile_by_id(0)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
# a C3-sized tile: 8 bands x 1024^2 uint16
rng = np.random.default_rng(0)
yy, xx = np.mgrid[0:1024, 0:2048]
big = np.stack([(3000 + 900 * np.sin(xx / 23.0 + b) * np.cos(yy / 17.0) + rng.integers(-20, 20, xx.shape)).astype(np.uint16) for b in range(8)])
write_geotiff(d / "big.tif", big, (10.0, 0.0, 3e5, 0.0, -10.0, 4e6), "EPSG:32633", None)
SpatialFLACEncoder(tile_size=1024).encode(d / "big.tif", d / "big.flac", streaming=True)
sb = SpatialFLACStreamer(d / "big.flac")
for _ in range(3): sb.get_tile_by_id(1)
ts = []
for _ in range(15):
    t0 = time.perf_counter(); sb.get_tile_by_id(1); ts.append(time.perf_counter() - t0)
print("C3-size tile get_tile_by_id ms: median %.3f min %.3f" % (1e3 * np.median(ts), 1e3 * min(ts)))
# both tiles of that container through get_tiles_by_bbox (multi-band tiles of one size: views of a (tile, band, h, w) block)
for _ in range(3): sb.get_tiles_by_bbox(-1e12, -1e12, 1e12, 1e12)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); res = sb.get_tiles_by_bbox(-1e12, -1e12, 1e12, 1e12); ts.append(time.perf_counter() - t0)
print("2 x C3-size tiles get_tiles_by_bbox ms: median %.3f min %.3f (%d tiles, %.1f MB of pixels)" % (1e3 * np.median(ts), 1e3 * min(ts), len(res), sum(a.nbytes for a, _ in res) / 1e6))
assert all(np.array_equal(a, big[:, m["window"]["row_off"]:m["window"]["row_off"] + 1024, m["window"]["col_off"]:m["window"]["col_off"] + 1024]) for a, m in res)
