"""BASELINE configs 1 and 2 (the reference's own CPU-runnable files) through the public API: wall clock per call on the GPU
path, and the CPU oracle port (1 core) on the same samples.  Prints one JSON object.  Not a bench line: these files are 0.26 M
and 0.2 M samples, the calls are launch/latency bound."""
import json, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flac_raster_b200 import RasterFLACConverter, SpatialFLACEncoder, SpatialFLACStreamer
from flac_raster_b200.tiffio import read_geotiff
from oracle import flac_oracle as fo, normalization_oracle as no

G = ROOT / "tests" / "golden"


def med(fn, n=7):
    fn(); fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


out = {}
with tempfile.TemporaryDirectory() as d:
    d = Path(d)
    conv = RasterFLACConverter()
    dem = read_geotiff(G / "sample_dem.tif").data
    n1 = int(dem.size)
    t_enc = med(lambda: conv.tiff_to_flac(G / "sample_dem.tif", d / "dem.flac", compression_level=5))
    t_dec = med(lambda: conv.flac_to_tiff(d / "dem.flac", d / "dem.tif"))
    a, prm = no.normalize_to_audio(dem.reshape(-1, 1), 16)
    t_oe = med(lambda: fo.encode(a, 16, 44100, 5), 3)
    blob, _ = fo.encode(a, 16, 44100, 5)
    t_od = med(lambda: fo.decode(blob), 3)
    out["c1"] = {"samples": n1, "gpu_tiff_to_flac_ms": 1e3 * t_enc, "gpu_flac_to_tiff_ms": 1e3 * t_dec,
                 "cpu_port_encode_ms": 1e3 * t_oe, "cpu_port_decode_ms": 1e3 * t_od, "flac_bytes": (d / "dem.flac").stat().st_size,
                 "note": "GPU calls include TIFF read/write and file I/O; CPU rows are the codec only"}
    rgb = read_geotiff(G / "sample_rgb.tif").data
    n2 = int(rgb.size)
    enc = SpatialFLACEncoder(tile_size=512)
    t_enc2 = med(lambda: enc.encode(G / "sample_rgb.tif", d / "rgb.flac", streaming=True))
    st = SpatialFLACStreamer(d / "rgb.flac")
    t_get = med(lambda: st.get_tile_by_id(0))
    a2, _ = no.normalize_to_audio(rgb.reshape(3, -1).T, 16)
    t_oe2 = med(lambda: fo.encode(a2, 16, 44100, 5), 3)
    blob2, fs2 = fo.encode(a2, 16, 44100, 5)
    t_od2 = med(lambda: fo.decode(blob2), 3)
    out["c2"] = {"samples": n2, "gpu_encode_streaming_ms": 1e3 * t_enc2, "gpu_get_tile_by_id_ms": 1e3 * t_get,
                 "cpu_port_encode_ms": 1e3 * t_oe2, "cpu_port_decode_ms": 1e3 * t_od2, "frame_bytes": int(fs2.sum()),
                 "libflac_golden_frame_bytes": 178857}
print(json.dumps(out))
