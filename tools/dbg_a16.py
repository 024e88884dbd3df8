"""Debug: int16-audio vs int32-audio encode of the golden DEM, frame by frame (which frames differ, subframe type byte)."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from flac_raster_b200.engine import Engine, tile_grid
from flac_raster_b200.tiffio import read_geotiff
from flac_raster_b200.normalization import sample_rates_for_pixel_counts
from oracle import flac_oracle as fo, normalization_oracle as no

a = read_geotiff(ROOT / "tests/golden/sample_dem.tif").data
if a.ndim == 2:
    a = a[None]
eng = Engine(0)
dev = torch.from_numpy(a.copy()).cuda()
tiles = tile_grid(a.shape[1], a.shape[2], max(a.shape[1:]))
a32, base, npx, mm, bits = eng.normalize_tiles(dev, tiles)
total = int(npx.sum()) * a.shape[0]
a32 = a32[: total * 4].view(torch.int32).clone()
a16, *_ = eng.normalize_tiles(dev, tiles, audio_i16=True)
print("samples equal:", torch.equal(a16.to(torch.int32), a32), a16.dtype, a16.data_ptr() % 16, a32.data_ptr() % 16)
rates = sample_rates_for_pixel_counts(npx)
want = a32.cpu().numpy().reshape(a.shape[0], -1).T
oenc, ofs = fo.encode(np.ascontiguousarray(want), 16, int(rates[0]), 5)
oframes = oenc[len(oenc) - int(ofs.sum()):]
res = {}
for name, aud in (("a32", a32), ("a16", a16)):
    p, o, s, fb, sb = eng.encode_audio(aud, npx, base, rates, a.shape[0], 16, 5, 4096, payload_name=name)
    res[name] = (p.cpu().numpy().tobytes(), fb.cpu().numpy())
    print(name, "bytes", len(res[name][0]), "== oracle:", res[name][0] == oframes)
fo_off = np.concatenate([[0], np.cumsum(ofs)])
for name in res:
    blob, fb = res[name]
    off = np.concatenate([[0], np.cumsum(fb)])
    bad = [k for k in range(len(fb)) if blob[off[k]:off[k + 1]] != oframes[fo_off[k]:fo_off[k + 1]]]
    print(name, "frames differing from oracle:", bad[:20], "of", len(fb))
    for k in bad[:4]:
        print("  frame", k, "size", fb[k], "oracle", ofs[k], "type byte", hex(blob[off[k] + 6]), "oracle", hex(oframes[fo_off[k] + 6]))
