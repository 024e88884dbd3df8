"""FLAC container plumbing kept on the host: metadata blocks and VORBIS_COMMENT tags.

Frames (all codec arithmetic) are produced/consumed by the CUDA engine; this
module only writes/reads the byte-level metadata the reference obtains from
libFLAC's stream header (docs/sonos-pyflac.txt:2200-2212) and from mutagen
(reference converter.py:263-390).  No mutagen dependency: VORBIS_COMMENT is a
little-endian length-prefixed list (RFC 9639 section 8.6).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

VENDOR = "flac-raster-b200 0.1 (sm_100a CUDA FLAC engine)"

BLOCK_STREAMINFO = 0
BLOCK_PADDING = 1
BLOCK_VORBIS_COMMENT = 4


@dataclass
class StreamInfo:
    min_blocksize: int
    max_blocksize: int
    min_framesize: int
    max_framesize: int
    sample_rate: int
    channels: int
    bits_per_sample: int
    total_samples: int
    md5: bytes = b"\0" * 16

    def pack(self) -> bytes:
        v = ((self.sample_rate & 0xFFFFF) << 44) | ((self.channels - 1) << 41) | \
            ((self.bits_per_sample - 1) << 36) | (self.total_samples & 0xFFFFFFFFF)
        return (struct.pack(">HH", self.min_blocksize, self.max_blocksize)
                + self.min_framesize.to_bytes(3, "big") + self.max_framesize.to_bytes(3, "big")
                + v.to_bytes(8, "big") + self.md5)

    @classmethod
    def unpack(cls, b: bytes) -> "StreamInfo":
        if len(b) < 34:
            raise ValueError("STREAMINFO too short")
        mn, mx = struct.unpack(">HH", b[:4])
        v = int.from_bytes(b[10:18], "big")
        return cls(mn, mx, int.from_bytes(b[4:7], "big"), int.from_bytes(b[7:10], "big"),
                   v >> 44, ((v >> 41) & 7) + 1, ((v >> 36) & 31) + 1, v & 0xFFFFFFFFF, bytes(b[18:34]))


@dataclass
class FlacHeader:
    streaminfo: StreamInfo
    tags: Dict[str, List[str]] = field(default_factory=dict)
    vendor: str = ""
    first_frame_offset: int = 0


def _block(block_type: int, body: bytes, last: bool) -> bytes:
    return bytes([(0x80 if last else 0) | block_type]) + len(body).to_bytes(3, "big") + body


def pack_vorbis_comment(tags: Dict[str, str] | List[Tuple[str, str]], vendor: str = VENDOR) -> bytes:
    items = list(tags.items()) if isinstance(tags, dict) else list(tags)
    v = vendor.encode("utf-8")
    out = [struct.pack("<I", len(v)), v, struct.pack("<I", len(items))]
    for k, val in items:
        e = f"{k}={val}".encode("utf-8")
        out += [struct.pack("<I", len(e)), e]
    return b"".join(out)


def unpack_vorbis_comment(body: bytes) -> Tuple[str, Dict[str, List[str]]]:
    pos = 0
    (vl,) = struct.unpack_from("<I", body, pos); pos += 4
    vendor = body[pos:pos + vl].decode("utf-8", "replace"); pos += vl
    (n,) = struct.unpack_from("<I", body, pos); pos += 4
    tags: Dict[str, List[str]] = {}
    for _ in range(n):
        (el,) = struct.unpack_from("<I", body, pos); pos += 4
        e = body[pos:pos + el].decode("utf-8", "replace"); pos += el
        if "=" in e:
            k, val = e.split("=", 1)
            tags.setdefault(k.upper(), []).append(val)
    return vendor, tags


def build_header(si: StreamInfo, tags: Optional[Dict[str, str]] = None, vendor: str = VENDOR,
                 padding: int = 0) -> bytes:
    """'fLaC' + STREAMINFO + VORBIS_COMMENT [+ PADDING]."""
    vc = pack_vorbis_comment(tags or {}, vendor)
    out = b"fLaC" + _block(BLOCK_STREAMINFO, si.pack(), False) + _block(BLOCK_VORBIS_COMMENT, vc, padding <= 0)
    if padding > 0:
        out += _block(BLOCK_PADDING, b"\0" * padding, True)
    return out


def parse_header(data: bytes | memoryview) -> FlacHeader:
    """Walk the metadata chain; returns STREAMINFO, tags and the first frame offset."""
    if len(data) < 8 or bytes(data[:4]) != b"fLaC":
        raise ValueError("not a FLAC stream (missing fLaC marker)")
    pos, last, si, tags, vendor = 4, False, None, {}, ""
    while not last:
        if pos + 4 > len(data):
            raise ValueError("truncated metadata")
        h = data[pos]
        last = bool(h & 0x80)
        btype = h & 0x7F
        blen = int.from_bytes(bytes(data[pos + 1:pos + 4]), "big")
        pos += 4
        if pos + blen > len(data):
            raise ValueError("truncated metadata block")
        body = bytes(data[pos:pos + blen])
        if btype == BLOCK_STREAMINFO:
            si = StreamInfo.unpack(body)
        elif btype == BLOCK_VORBIS_COMMENT:
            vendor, tags = unpack_vorbis_comment(body)
        elif btype == 127:
            raise ValueError("invalid metadata block type")
        pos += blen
    if si is None:
        raise ValueError("missing STREAMINFO")
    return FlacHeader(si, tags, vendor, pos)
