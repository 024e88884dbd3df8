#!/usr/bin/env python
"""Prototype: does running chunks of tiles on two CUDA streams overlap the HBM-bound kernels (min/max, normalise,
emit, CRC, denormalise) with the issue-bound ones (stats, code, skim, decode)?  Prints ms per C3 step for K chunks."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from flac_raster_b200 import _native as nat, synth  # noqa: E402
from flac_raster_b200.engine import Engine, tile_grid  # noqa: E402
from flac_raster_b200.normalization import audio_params_for  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
L = nat.lib()
div = int(sys.argv[1]) if len(sys.argv) > 1 else 1
side = 10980 // div
raster = synth.sentinel2_like(side, side, 8, device=dev)
tiles = tile_grid(side, side, 1024 // div if div > 1 else 1024)
nb = 8
level = 5
samples = raster.numel()


def timeit(fn, n=5, w=2):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


class Lane:
    def __init__(self, prio):
        self.stream = torch.cuda.Stream(priority=prio)
        self.eng = Engine(dev)


def chunks_of(n, k):
    b = np.linspace(0, n, k + 1).astype(int)
    return [(int(b[i]), int(b[i + 1])) for i in range(k) if b[i + 1] > b[i]]


def encode_chunked(K, lanes):
    parts = []
    cur = torch.cuda.current_stream()
    for ln in lanes:
        ln.stream.wait_stream(cur)
    pend = []
    ch = chunks_of(len(tiles), K)
    norm = {}

    def do_norm(j):
        i0, i1 = ch[j]
        ln = lanes[j % len(lanes)]
        with torch.cuda.stream(ln.stream):
            norm[j] = ln.eng.normalize_tiles(raster, tiles[i0:i1])

    do_norm(0)
    for j, (i0, i1) in enumerate(ch):
        if j + 1 < len(ch):
            do_norm(j + 1)
        ln = lanes[j % len(lanes)]
        with torch.cuda.stream(ln.stream):
            audio, base, npx, d_mm, bits = norm.pop(j)
            rates = np.array([audio_params_for((int(t["h"]), int(t["w"])), "uint16")[0] for t in tiles[i0:i1]], dtype=np.uint32)
            parts.append(ln.eng.encode_audio(audio, npx, base, rates, nb, 16, level, 4096, payload_name=f"payload{j}"))
    for ln in lanes:
        cur.wait_stream(ln.stream)
    return parts


base_eng = Engine(dev)
t_base = timeit(lambda: base_eng.encode_tiles(raster, tiles, level))
print(f"encode baseline: {t_base:.2f} ms  {samples / t_base / 1e6:.1f} GS/s")
enc = base_eng.encode_tiles(raster, tiles, level)
for prio in ((0, 0), (0, -1)):
    lanes = [Lane(prio[0]), Lane(prio[1])]
    for K in (2, 4, 6, 8, 12):
        t = timeit(lambda: encode_chunked(K, lanes))
        print(f"encode 2 lanes prio={prio} K={K}: {t:.2f} ms  {samples / t / 1e6:.1f} GS/s")
    del lanes

# ---------------------------------------------------------------- decode
payload = torch.cat([enc.payload, torch.zeros(64, dtype=torch.uint8, device=dev)])
out = torch.zeros_like(raster)


def decode_base():
    audio, base, st = base_eng.decode_streams(payload, enc.offsets, enc.sizes, enc.n_samples, enc.sample_rates, nb, 16, 4096)
    base_eng.denormalize_tiles(audio, base, tiles, enc.minmax, 32767.0, out)


t_base = timeit(decode_base)
print(f"decode baseline: {t_base:.2f} ms  {samples / t_base / 1e6:.1f} GS/s")
assert torch.equal(out.view(torch.int16), raster.view(torch.int16))


def decode_chunked(K, lanes):
    cur = torch.cuda.current_stream()
    for ln in lanes:
        ln.stream.wait_stream(cur)
    keep = []
    for j, (i0, i1) in enumerate(chunks_of(len(tiles), K)):
        ln = lanes[j % len(lanes)]
        with torch.cuda.stream(ln.stream):
            audio, base, st = ln.eng.decode_streams(payload, enc.offsets[i0:i1], enc.sizes[i0:i1], enc.n_samples[i0:i1],
                                                    enc.sample_rates[i0:i1], nb, 16, 4096, sync=False, name=f"dec{j // len(lanes) % 2}")
            ln.eng.denormalize_tiles(audio, base, tiles[i0:i1], enc.minmax[i0:i1], 32767.0, out, sync=False)
            keep.append(audio)
    for ln in lanes:
        cur.wait_stream(ln.stream)


for prio in ((0, 0), (0, -1)):
    lanes = [Lane(prio[0]), Lane(prio[1])]
    for K in (2, 4, 6, 8, 12):
        out.zero_()
        t = timeit(lambda: decode_chunked(K, lanes))
        ok = torch.equal(out.view(torch.int16), raster.view(torch.int16))
        print(f"decode 2 lanes prio={prio} K={K}: {t:.2f} ms  {samples / t / 1e6:.1f} GS/s ok={ok}")
    del lanes
lanes = [Lane(0), Lane(0), Lane(0)]
for K in (6, 12):
    t = timeit(lambda: decode_chunked(K, lanes))
    print(f"decode 3 lanes K={K}: {t:.2f} ms  {samples / t / 1e6:.1f} GS/s")
    t = timeit(lambda: encode_chunked(K, lanes))
    print(f"encode 3 lanes K={K}: {t:.2f} ms  {samples / t / 1e6:.1f} GS/s")
