"""ctypes binding of oracle/flac_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(flac_raster_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libflac_oracle.so"

MAX_LPC = 32
DESC_PARAMS = 64


class StreamInfo(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bps", C.c_uint32),
        ("min_blocksize", C.c_uint32), ("max_blocksize", C.c_uint32),
        ("min_framesize", C.c_uint32), ("max_framesize", C.c_uint32),
        ("total_samples_streaminfo", C.c_uint64),
        ("samples_decoded", C.c_uint64),
        ("n_frames", C.c_uint32), ("first_frame_offset", C.c_uint32),
        ("crc8_errors", C.c_uint32), ("crc16_errors", C.c_uint32),
        ("bytes_consumed", C.c_uint64),
        ("md5", C.c_uint8 * 16),
        ("vorbis_offset", C.c_uint32), ("vorbis_length", C.c_uint32),
    ]


class SubframeDesc(C.Structure):
    _fields_ = [
        ("frame", C.c_uint32), ("channel", C.c_uint32), ("type", C.c_uint32),
        ("order", C.c_uint32), ("wasted", C.c_uint32), ("precision", C.c_uint32),
        ("shift", C.c_int32), ("coefs", C.c_int32 * MAX_LPC),
        ("method", C.c_uint32), ("partition_order", C.c_uint32),
        ("params", C.c_uint32 * DESC_PARAMS), ("n_escape", C.c_uint32),
        ("nbits", C.c_uint32), ("blocksize", C.c_uint32), ("ch_assign", C.c_uint32),
        ("frame_bytes", C.c_uint32), ("frame_offset", C.c_uint32),
    ]

    def as_dict(self):
        n = self.order
        return dict(frame=self.frame, channel=self.channel, type=self.type, order=n,
                    wasted=self.wasted, precision=self.precision, shift=self.shift,
                    coefs=list(self.coefs[:n]), method=self.method,
                    partition_order=self.partition_order,
                    params=list(self.params[:min(DESC_PARAMS, 1 << self.partition_order)]),
                    n_escape=self.n_escape, nbits=self.nbits, blocksize=self.blocksize,
                    ch_assign=self.ch_assign, frame_bytes=self.frame_bytes,
                    frame_offset=self.frame_offset)


def build(force: bool = False) -> Path:
    src = _HERE / "flac_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", str(_SO), str(src), "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        L.fo_decode_stream.restype = C.c_int
        L.fo_decode_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                       C.POINTER(StreamInfo), C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_size_t)]
        L.fo_encode_stream.restype = C.c_size_t
        L.fo_encode_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.c_uint32, C.c_int, C.c_char_p,
                                       C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_size_t]
        L.fo_crc8.restype = C.c_uint8
        L.fo_crc8.argtypes = [C.c_void_p, C.c_size_t]
        L.fo_crc16.restype = C.c_uint16
        L.fo_crc16.argtypes = [C.c_void_p, C.c_size_t]
        L.fo_sizeof_info.restype = C.c_size_t
        L.fo_sizeof_desc.restype = C.c_size_t
        assert L.fo_sizeof_info() == C.sizeof(StreamInfo)
        assert L.fo_sizeof_desc() == C.sizeof(SubframeDesc)
        _lib = L
    return _lib


def decode(data: bytes, want_descs: bool = False, max_samples: int | None = None):
    """Decode one FLAC stream.  Returns (samples (N,C) int32, StreamInfo[, descs])."""
    L = lib()
    buf = np.frombuffer(data, dtype=np.uint8)
    info = StreamInfo()
    # first pass without output to learn the size
    nd = C.c_size_t(0)
    rc = L.fo_decode_stream(buf.ctypes.data, buf.size, None, 0, C.byref(info), None, 0, C.byref(nd))
    if rc != 0:
        raise ValueError(f"oracle decode failed rc={rc}")
    n, ch = int(info.samples_decoded), int(info.channels)
    out = np.empty((n, ch), dtype=np.int32)
    ndesc = int(info.n_frames) * ch if want_descs else 0
    descs = (SubframeDesc * max(ndesc, 1))()
    rc = L.fo_decode_stream(buf.ctypes.data, buf.size, out.ctypes.data, out.size, C.byref(info),
                            C.cast(descs, C.c_void_p) if want_descs else None, ndesc, C.byref(nd))
    if rc != 0:
        raise ValueError(f"oracle decode failed rc={rc}")
    if want_descs:
        return out, info, [descs[i].as_dict() for i in range(nd.value)]
    return out, info


def encode(samples: np.ndarray, bps: int, sample_rate: int, level: int = 5, blocksize: int = 4096,
           finalize: bool = False, vendor: bytes | None = None, want_descs: bool = False, mid_side: bool = True):
    """Encode (N,C) int samples into a full FLAC stream (bytes).  mid_side=False codes two-channel input as
    independent channels (what the GPU encoder emits today); True follows libFLAC's presets."""
    L = lib()
    a = np.ascontiguousarray(samples).astype(np.int32)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    n, ch = a.shape
    cap = 8192 + n * ch * 5 + (n // blocksize + 2) * 64
    out = np.empty(cap, dtype=np.uint8)
    nframes = (n + blocksize - 1) // blocksize
    fs = np.zeros(max(nframes, 1), dtype=np.uint32)
    ndesc = nframes * ch if want_descs else 0
    descs = (SubframeDesc * max(ndesc, 1))()
    sz = L.fo_encode_stream(a.ctypes.data, n, ch, bps, sample_rate, level, blocksize, int(bool(finalize)) | (0 if mid_side else 2),
                            vendor, out.ctypes.data, cap, fs.ctypes.data, fs.size,
                            C.cast(descs, C.c_void_p) if want_descs else None, ndesc)
    if sz == 0:
        raise ValueError("oracle encode failed")
    data = out[:sz].tobytes()
    if want_descs:
        return data, fs[:nframes], [descs[i].as_dict() for i in range(ndesc)]
    return data, fs[:nframes]


def crc8(b: bytes) -> int:
    a = np.frombuffer(b, dtype=np.uint8)
    return int(lib().fo_crc8(a.ctypes.data, a.size))


def crc16(b: bytes) -> int:
    a = np.frombuffer(b, dtype=np.uint8)
    return int(lib().fo_crc16(a.ctypes.data, a.size))
