"""BASELINE config 5 through the public API: grouped (pipelined) against one-batch decode, read piece size / thread sweeps."""
import sys, time, os, tempfile
sys.path.insert(0, '.')
import numpy as np, torch
from flac_raster_b200 import synth
from flac_raster_b200.spatial_encoder import SpatialFLACStreamer, build_streaming_container, write_streaming_container
side_tiles = 64
H = W = side_tiles * 512
dev = torch.device('cuda', 0)
strip = synth.dem_int16_tiles(side_tiles * side_tiles, 512, device=dev)
raster = strip.reshape(side_tiles, side_tiles, 512, 512).permute(0, 2, 1, 3).reshape(1, H, W).contiguous()
index, headers, enc = build_streaming_container(raster, (1.0, 0.0, 0.0, 0.0, -1.0, float(H)), "EPSG:32633", None, "int16", 512, 5)
payload = enc.payload.cpu().numpy()
path = os.path.join(tempfile.gettempdir(), "c5.flac")
write_streaming_container(path, index, headers, payload, enc.offsets, enc.sizes)
del raster, strip, enc, payload
st = SpatialFLACStreamer(path)
print("cpus", os.cpu_count())


def run(tag, n=4):
    ts = []
    for rep in range(n):
        t1 = time.perf_counter(); res = st.get_tiles_by_bbox(0.0, 0.0, float(W), float(H)); ts.append(time.perf_counter() - t1); del res
    print(tag, [round(1e3 * t, 1) for t in ts])


run("warm grouped")
os.environ["FRB_NO_GROUPED_DECODE"] = "1"; run("warm one")
for piece, thr in ((32, 8), (8, 8), (8, 16), (4, 16), (16, 16)):
    os.environ["FRB_READ_PIECE_MB"] = str(piece); os.environ["FRB_READ_THREADS"] = str(thr)
    os.environ["FRB_NO_GROUPED_DECODE"] = "1"; run(f"one     piece {piece} MB threads {thr}", 3)
    os.environ.pop("FRB_NO_GROUPED_DECODE"); run(f"grouped piece {piece} MB threads {thr}", 3)
# bare read of the file into the pinned staging buffer (no GPU work): the floor of the read stage
import torch
from concurrent.futures import ThreadPoolExecutor
size = os.path.getsize(path)
pin = torch.empty(size, dtype=torch.uint8, pin_memory=True)
view = memoryview(pin.numpy())
fd = os.open(path, os.O_RDONLY)
for piece, thr in ((32, 8), (8, 16), (4, 32)):
    jobs = [(o, min(piece << 20, size - o)) for o in range(0, size, piece << 20)]
    def rd(j):
        o, n = j; got = 0
        while got < n:
            got += os.preadv(fd, [view[o + got:o + n]], o + got)
    for rep in range(2):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(thr) as ex: list(ex.map(rd, jobs))
        dt = time.perf_counter() - t0
    print(f"bare read piece {piece} MB threads {thr}: {1e3*dt:.1f} ms = {size/dt/1e9:.1f} GB/s")
