"""BASELINE config 5: get_tiles_by_bbox over 4096 tiles of 512x512 int16 -- end-to-end timing of the public API."""
import sys, time, cProfile, pstats, os, tempfile
sys.path.insert(0, '.')
import numpy as np, torch
from flac_raster_b200 import synth
from flac_raster_b200.spatial_encoder import SpatialFLACStreamer, build_streaming_container, write_streaming_container
side_tiles = int(os.environ.get("SIDE_TILES", "64"))
H = W = side_tiles * 512
dev = torch.device('cuda', 0)
strip = synth.dem_int16_tiles(side_tiles * side_tiles, 512, device=dev)          # (1, n*512, 512)
raster = strip.reshape(side_tiles, side_tiles, 512, 512).permute(0, 2, 1, 3).reshape(1, H, W).contiguous()
t0 = time.perf_counter()
index, headers, enc = build_streaming_container(raster, (1.0, 0.0, 0.0, 0.0, -1.0, float(H)), "EPSG:32633", None, "int16", 512, 5)
payload = enc.payload.cpu().numpy()
t1 = time.perf_counter()
path = os.path.join(tempfile.gettempdir(), "c5.flac")
write_streaming_container(path, index, headers, payload, enc.offsets, enc.sizes)
t2 = time.perf_counter()
print(f"encode+container {t1-t0:.3f}s write {t2-t1:.3f}s size {os.path.getsize(path)/1e6:.1f} MB tiles {len(headers)}")
host = raster.cpu().numpy()
for rep in range(2):
    t0 = time.perf_counter()
    st = SpatialFLACStreamer(path)
    t1 = time.perf_counter()
    res = st.get_tiles_by_bbox(0.0, 0.0, float(W), float(H))
    t2 = time.perf_counter()
    print(f"open {t1-t0:.3f}s  get_tiles_by_bbox {t2-t1:.3f}s  -> {len(res)} tiles, {sum(a.size for a, _ in res)/ (t2-t1)/1e9:.3f} GSamples/s")
ok = all(np.array_equal(a, host[:, m['window']['row_off']:m['window']['row_off']+512, m['window']['col_off']:m['window']['col_off']+512]) for a, m in res[:64])
print("first 64 tiles identical:", ok)
pr = cProfile.Profile(); pr.enable(); res = st.get_tiles_by_bbox(0.0, 0.0, float(W), float(H)); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
