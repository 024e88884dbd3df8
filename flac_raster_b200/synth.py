"""Deterministic synthetic rasters generated on the device (benchmarks and full-size tests).

Shapes and formulas follow SURVEY.md section 8(d): a per-pixel counter hash
h = splitmix64(seed ^ (band<<40 | y<<20 | x)), u = (h>>11) * 2^-53, plus smooth terrain terms.
"""
from __future__ import annotations

import torch


def _splitmix64(x: torch.Tensor) -> torch.Tensor:
    # int64 arithmetic wraps, which is exactly the uint64 arithmetic splitmix64 needs
    def lsr(v, n):
        return (v >> n) & ((1 << (64 - n)) - 1)
    x = x + (-7046029254386353131)                       # 0x9E3779B97F4A7C15
    x = (x ^ lsr(x, 30)) * (-4658895280553007687)        # 0xBF58476D1CE4E5B9
    x = (x ^ lsr(x, 27)) * (-7723592293110705685)        # 0x94D049BB133111EB
    return x ^ lsr(x, 31)


def _uniform(seed: int, band: int, ys: torch.Tensor, xs: torch.Tensor, salt: int = 0) -> torch.Tensor:
    key = (band << 40) | (ys.to(torch.int64) << 20) | xs.to(torch.int64)
    h = _splitmix64(key ^ (seed + salt * 0x632BE5AB))
    return ((h >> 11) & ((1 << 53) - 1)).to(torch.float64) * (2.0 ** -53)


def sentinel2_like(height: int, width: int, bands: int = 8, seed: int = 0, device="cuda", row0: int = 0) -> torch.Tensor:
    """C3: uint16 reflectance-like bands clipped to [0, 11672]; returned as a (bands,H,W) uint16 tensor."""
    out = torch.empty((bands, height, width), dtype=torch.int16, device=device)
    xs = torch.arange(width, device=device, dtype=torch.float64)[None, :]
    rows_per = max(1, (1 << 24) // max(width, 1))
    for b in range(bands):
        for r in range(0, height, rows_per):
            ys = torch.arange(row0 + r, row0 + min(height, r + rows_per), device=device, dtype=torch.float64)[:, None]
            yi, xi = ys.expand(-1, width), xs.expand(ys.shape[0], -1)
            u = _uniform(seed, b, yi, xi, 0) + _uniform(seed, b, yi, xi, 1) + _uniform(seed, b, yi, xi, 2)
            v = 1800 + 300 * b + 1400 * torch.sin(xi / 97 + b) * torch.cos(yi / 131) + 500 * torch.sin((xi + yi) / 23) + 120 * (u - 1.5)
            v = v.clamp_(0, 11672).to(torch.int32)
            out[b, r:r + ys.shape[0]] = v.to(torch.int16)          # values <= 11672 fit; reinterpret below
    return out.view(torch.uint16)


def dem_float32(height: int, width: int, seed: int = 0, device="cuda") -> torch.Tensor:
    """C4: float32 terrain, (1,H,W)."""
    out = torch.empty((1, height, width), dtype=torch.float32, device=device)
    xs = torch.arange(width, device=device, dtype=torch.float64)[None, :]
    rows_per = max(1, (1 << 24) // max(width, 1))
    for r in range(0, height, rows_per):
        ys = torch.arange(r, min(height, r + rows_per), device=device, dtype=torch.float64)[:, None]
        yi, xi = ys.expand(-1, width), xs.expand(ys.shape[0], -1)
        u = _uniform(seed, 0, yi, xi)
        z = 1000 + 300 * torch.sin(xi / 300) * torch.cos(yi / 200) + 150 * torch.sin(xi / 37) * torch.sin(yi / 41) + 2 * (u - 0.5)
        out[0, r:r + ys.shape[0]] = z.to(torch.float32)
    return out


def dem_int16_tiles(n_tiles: int, tile: int = 512, seed: int = 0, device="cuda", first_tile: int = 0) -> torch.Tensor:
    """C5: n_tiles stacked vertically as one (1, n_tiles*tile, tile) int16 raster; tile index in the phase."""
    out = torch.empty((1, n_tiles * tile, tile), dtype=torch.int16, device=device)
    xs = torch.arange(tile, device=device, dtype=torch.float64)[None, :]
    group = max(1, (1 << 24) // (tile * tile))
    for t0 in range(0, n_tiles, group):
        nt = min(group, n_tiles - t0)
        ys = torch.arange(nt * tile, device=device, dtype=torch.float64)[:, None]
        tid = torch.div(ys, tile, rounding_mode="floor") + (first_tile + t0)
        yl = ys - (tid - first_tile - t0) * tile
        yi, xi = yl.expand(-1, tile), xs.expand(ys.shape[0], -1)
        u = _uniform(seed, 0, (tid * tile + yl).expand(-1, tile), xi)
        z = 3000 * torch.sin(xi / 61 + tid * 0.37) * torch.cos(yi / 47 + tid * 0.11) + 50 * (2 * u - 1)
        out[0, t0 * tile:(t0 + nt) * tile] = z.to(torch.int16)
    return out
