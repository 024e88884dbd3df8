"""Minimal GeoTIFF reader/writer (numpy only) used when rasterio is not installed.

Raster I/O is outside the accelerated path (the reference uses rasterio/GDAL,
converter.py:73-79, :253-257); this module exists so the drop-in API works in
environments without GDAL.  Reads baseline TIFF (strips or tiles, uncompressed
or Deflate/LZW-free... i.e. compression 1 or 8/32946), chunky or planar, 8/16/32/64-bit
unsigned, signed and float samples, plus the GeoTIFF tags needed for the affine
transform and CRS (ModelPixelScale 33550, ModelTiepoint 33922,
ModelTransformation 34264, GeoKeyDirectory 34735, GDAL_NODATA 42113).
Writes uncompressed, band-interleaved-by-pixel (chunky) strips with the same tags.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d", 16: "Q", 17: "q"}
_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8}


@dataclass
class Raster:
    """(bands, H, W) array plus the metadata the converter needs (rasterio's src.meta subset)."""
    data: np.ndarray
    transform: Optional[Tuple[float, float, float, float, float, float]] = None   # (a, b, c, d, e, f)
    crs: Optional[str] = None
    nodata: Optional[float] = None
    geokeys: Optional[Tuple[int, ...]] = None
    geo_doubles: Optional[Tuple[float, ...]] = None
    geo_ascii: Optional[str] = None

    @property
    def count(self) -> int:
        return self.data.shape[0]

    @property
    def height(self) -> int:
        return self.data.shape[1]

    @property
    def width(self) -> int:
        return self.data.shape[2]

    @property
    def bounds(self) -> Dict[str, float]:
        t = self.transform or (1.0, 0.0, 0.0, 0.0, -1.0, 0.0)
        left, top = t[2], t[5]
        right = left + self.width * t[0]
        bottom = top + self.height * t[4]
        return {"left": left, "bottom": bottom, "right": right, "top": top}


def window_transform(transform, col_off: int, row_off: int):
    """Affine of a window (rasterio src.window_transform; cli.py:561)."""
    a, b, c, d, e, f = transform
    return (a, b, c + col_off * a + row_off * b, d, e, f + col_off * d + row_off * e)


def _read_ifd(buf: bytes, off: int, bo: str, big: bool):
    tags = {}
    if big:
        (n,) = struct.unpack_from(bo + "Q", buf, off); off += 8
        esz = 20
    else:
        (n,) = struct.unpack_from(bo + "H", buf, off); off += 2
        esz = 12
    for i in range(n):
        e = off + i * esz
        tag, typ = struct.unpack_from(bo + "HH", buf, e)
        if big:
            (cnt,) = struct.unpack_from(bo + "Q", buf, e + 4)
            voff, inline = e + 12, 8
        else:
            (cnt,) = struct.unpack_from(bo + "I", buf, e + 4)
            voff, inline = e + 8, 4
        size = _TYPE_SIZE.get(typ, 1) * cnt
        if size > inline:
            (voff,) = struct.unpack_from(bo + ("Q" if big else "I"), buf, voff)
        raw = buf[voff:voff + size]
        if typ == 2:
            val = raw.split(b"\0")[0].decode("latin-1")
        elif typ in (5, 10):
            v = struct.unpack(bo + _TYPE_FMT[typ][0] * (2 * cnt), raw)
            val = tuple(v[2 * k] / v[2 * k + 1] if v[2 * k + 1] else 0.0 for k in range(cnt))
        else:
            val = struct.unpack(bo + _TYPE_FMT.get(typ, "B") * cnt, raw)
        tags[tag] = val
    return tags


def _crs_from_geokeys(tags) -> Optional[str]:
    gk = tags.get(34735)
    if not gk:
        return None
    n = gk[3]
    for k in range(n):
        key, loc, cnt, val = gk[4 + 4 * k: 8 + 4 * k]
        if key in (2048, 3072) and loc == 0 and val not in (0, 32767):
            return f"EPSG:{val}"
    return None


def read_geotiff(path) -> Raster:
    """Read the first image of a TIFF into a (bands, H, W) array."""
    try:
        import rasterio  # type: ignore

        with rasterio.open(path) as src:
            t = src.transform
            return Raster(src.read(), (t.a, t.b, t.c, t.d, t.e, t.f) if t else None,
                          src.crs.to_string() if src.crs else None, src.nodata)
    except ImportError:
        pass
    buf = Path(path).read_bytes()
    bo = "<" if buf[:2] == b"II" else ">"
    (magic,) = struct.unpack_from(bo + "H", buf, 2)
    if magic == 42:
        big = False
        (ifd,) = struct.unpack_from(bo + "I", buf, 4)
    elif magic == 43:
        big = True
        (ifd,) = struct.unpack_from(bo + "Q", buf, 8)
    else:
        raise ValueError(f"{path}: not a TIFF file")
    tags = _read_ifd(buf, ifd, bo, big)
    W, H = tags[256][0], tags[257][0]
    spp = tags.get(277, (1,))[0]
    bits = tags.get(258, (1,) * spp)
    fmt = tags.get(339, (1,) * spp)[0]
    comp = tags.get(259, (1,))[0]
    planar = tags.get(284, (1,))[0]
    pred = tags.get(317, (1,))[0]
    if len(set(bits)) != 1:
        raise ValueError("mixed bit depths are not supported")
    b = bits[0]
    kind = {1: "u", 2: "i", 3: "f"}.get(fmt, "u")
    dtype = np.dtype(f"{bo}{kind}{b // 8}")
    if comp not in (1, 8, 32946):
        raise ValueError(f"{path}: TIFF compression {comp} not supported without rasterio")

    def chunk(off, cnt):
        raw = buf[off:off + cnt]
        if comp != 1:
            raw = zlib.decompress(raw)
        return raw

    data = np.zeros((spp, H, W), dtype=dtype.newbyteorder("="))
    if 322 in tags:   # tiled
        tw, th = tags[322][0], tags[323][0]
        offs, cnts = tags[324], tags[325]
        tx, ty = (W + tw - 1) // tw, (H + th - 1) // th
        per_plane = tx * ty
        for idx, (o, c) in enumerate(zip(offs, cnts)):
            plane, t = divmod(idx, per_plane) if planar == 2 else (0, idx)
            r0, c0 = (t // tx) * th, (t % tx) * tw
            raw = chunk(o, c)
            if planar == 2:
                a = np.frombuffer(raw, dtype=dtype, count=tw * th).reshape(th, tw)
                if pred == 2:
                    a = np.cumsum(a, axis=1, dtype=a.dtype)
                data[plane, r0:r0 + th, c0:c0 + tw] = a[:min(th, H - r0), :min(tw, W - c0)]
            else:
                a = np.frombuffer(raw, dtype=dtype, count=tw * th * spp).reshape(th, tw, spp)
                if pred == 2:
                    a = np.cumsum(a, axis=1, dtype=a.dtype)
                data[:, r0:r0 + th, c0:c0 + tw] = a[:min(th, H - r0), :min(tw, W - c0)].transpose(2, 0, 1)
    else:
        rps = tags.get(278, (H,))[0]
        offs, cnts = tags[273], tags[279]
        spi = (H + rps - 1) // rps
        for idx, (o, c) in enumerate(zip(offs, cnts)):
            plane, sidx = divmod(idx, spi) if planar == 2 else (0, idx)
            r0 = sidx * rps
            rows = min(rps, H - r0)
            raw = chunk(o, c)
            if planar == 2:
                a = np.frombuffer(raw, dtype=dtype, count=rows * W).reshape(rows, W)
                if pred == 2:
                    a = np.cumsum(a, axis=1, dtype=a.dtype)
                data[plane, r0:r0 + rows] = a
            else:
                a = np.frombuffer(raw, dtype=dtype, count=rows * W * spp).reshape(rows, W, spp)
                if pred == 2:
                    a = np.cumsum(a, axis=1, dtype=a.dtype)
                data[:, r0:r0 + rows] = a.transpose(2, 0, 1)
    transform = None
    if 34264 in tags:
        m = tags[34264]
        transform = (m[0], m[1], m[3], m[4], m[5], m[7])
    elif 33550 in tags and 33922 in tags:
        sx, sy = tags[33550][0], tags[33550][1]
        tp = tags[33922]
        transform = (sx, 0.0, tp[3] - tp[0] * sx, 0.0, -sy, tp[4] + tp[1] * sy)
    nodata = None
    if 42113 in tags:
        try:
            nodata = float(tags[42113])
        except (TypeError, ValueError):
            nodata = None
    return Raster(data, transform, _crs_from_geokeys(tags), nodata, tags.get(34735), tags.get(34736), tags.get(34737))


def write_geotiff(path, data: np.ndarray, transform=None, crs: Optional[str] = None, nodata=None,
                  geokeys=None, geo_doubles=None, geo_ascii=None):
    """Write a (bands,H,W) or (H,W) array as an uncompressed little-endian GeoTIFF."""
    try:
        import rasterio  # type: ignore
        from rasterio.crs import CRS
        from rasterio.transform import Affine

        arr = data if data.ndim == 3 else data[None]
        meta = dict(driver="GTiff", width=arr.shape[2], height=arr.shape[1], count=arr.shape[0], dtype=arr.dtype, nodata=nodata)
        if crs:
            meta["crs"] = CRS.from_string(crs)
        if transform:
            meta["transform"] = Affine(*transform)
        with rasterio.open(path, "w", **meta) as dst:
            dst.write(arr)
        return
    except ImportError:
        pass
    arr = data if data.ndim == 3 else data[None]
    bands, H, W = arr.shape
    dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
    pix = np.ascontiguousarray(arr.transpose(1, 2, 0).astype(dt, copy=False))
    fmt = {"u": 1, "i": 2, "f": 3}[dt.kind]
    entries: List[Tuple[int, int, int, bytes]] = []   # tag, type, count, packed value

    def add(tag, typ, values):
        if typ == 2:
            raw = values.encode("latin-1") + b"\0"
            cnt = len(raw)
        else:
            values = tuple(values)
            cnt = len(values)
            raw = struct.pack("<" + _TYPE_FMT[typ] * cnt, *values)
        entries.append((tag, typ, cnt, raw))

    big = pix.nbytes > 0xF0000000
    add(256, 4, [W]); add(257, 4, [H]); add(258, 3, [dt.itemsize * 8] * bands); add(259, 3, [1])
    add(262, 3, [2 if bands == 3 and dt.itemsize == 1 else 1])
    add(277, 3, [bands]); add(278, 4, [H]); add(284, 3, [1]); add(339, 3, [fmt] * bands)
    if bands > 1 and not (bands == 3 and dt.itemsize == 1):
        add(338, 3, [0] * (bands - 1))
    if transform is not None:
        a, b_, c, d, e, f = transform
        if b_ == 0 and d == 0:
            add(33550, 12, [a, -e, 0.0]); add(33922, 12, [0.0, 0.0, 0.0, c, f, 0.0])
        else:
            add(34264, 12, [a, b_, 0.0, c, d, e, 0.0, f, 0, 0, 0, 0, 0, 0, 0, 1.0])
    if geokeys:
        add(34735, 3, geokeys)
        if geo_doubles:
            add(34736, 12, geo_doubles)
        if geo_ascii:
            add(34737, 2, geo_ascii)
    elif crs and crs.upper().startswith("EPSG:"):
        code = int(crs.split(":")[1])
        geographic = code in (4326, 4269, 4258, 4283) or 4000 <= code < 5000
        add(34735, 3, [1, 1, 0, 3, 1024, 0, 1, 2 if geographic else 1, 1025, 0, 1, 1,
                       2048 if geographic else 3072, 0, 1, code])
    if nodata is not None:
        add(42113, 2, repr(float(nodata)) if float(nodata) != int(float(nodata)) else str(int(float(nodata))))
    entries.sort(key=lambda e: e[0])
    with open(path, "wb") as fh:
        if big:
            hdr, esz, inline, ofmt = 16, 20, 8, "Q"
        else:
            hdr, esz, inline, ofmt = 8, 12, 4, "I"
        n = len(entries) + 2          # + StripOffsets, StripByteCounts
        ifd_size = (8 if big else 2) + n * esz + (8 if big else 4)
        extra_off = hdr + ifd_size
        extra = bytearray()
        packed = []
        all_entries = entries + [(273, 16 if big else 4, 1, None), (279, 16 if big else 4, 1, None)]
        all_entries.sort(key=lambda e: e[0])
        # first pass to size the out-of-line area
        sizes = sum((len(r) + 1) & ~1 for (_, _, _, r) in entries if r is not None and len(r) > inline)
        data_off = (extra_off + sizes + 15) & ~15
        for tag, typ, cnt, raw in all_entries:
            if tag == 273:
                raw = struct.pack("<" + ofmt, data_off)
            elif tag == 279:
                raw = struct.pack("<" + ofmt, pix.nbytes)
            if len(raw) > inline:
                val = struct.pack("<" + ofmt, extra_off + len(extra))
                extra += raw
                if len(extra) & 1:
                    extra += b"\0"
            else:
                val = raw.ljust(inline, b"\0")
            if big:
                packed.append(struct.pack("<HHQ", tag, typ, cnt) + val)
            else:
                packed.append(struct.pack("<HHI", tag, typ, cnt) + val)
        if big:
            fh.write(b"II" + struct.pack("<HHHQ", 43, 8, 0, 16))
            fh.write(struct.pack("<Q", n))
        else:
            fh.write(b"II" + struct.pack("<HI", 42, 8))
            fh.write(struct.pack("<H", n))
        fh.write(b"".join(packed))
        fh.write(struct.pack("<" + ofmt, 0))
        fh.write(bytes(extra))
        fh.write(b"\0" * (data_off - (extra_off + len(extra))))
        fh.write(pix.tobytes())
