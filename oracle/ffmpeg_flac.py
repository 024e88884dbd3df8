"""Independent production FLAC codec (FFmpeg libavcodec via ctypes: decoder as a checker, encoder + decoder as labelled
CPU baseline rows of bench.py) -- TEST INFRASTRUCTURE ONLY.

pyflac/libFLAC cannot be installed offline, but opencv-python-headless bundles FFmpeg 8
(libavformat 62 / libavcodec 62.11 / libavutil 60).  Decoding our streams with a second,
unrelated decoder checks "valid FLAC that another real decoder accepts" (SURVEY.md section 8c).
Struct offsets below were verified against this build (AVFormatContext.streams @48,
AVStream.codecpar @16, AVFrame.extended_data @96, nb_samples @112, format @116).
Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import tempfile

import numpy as np

_libs = None


def available() -> bool:
    try:
        _load()
        return True
    except Exception:  # noqa: BLE001
        return False


def _load():
    global _libs
    if _libs is not None:
        return _libs
    import cv2  # noqa: F401  (loads the bundled shared objects' dependencies)

    base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    def one(pat):
        m = sorted(glob.glob(os.path.join(base, pat)))
        if not m:
            raise ImportError(f"{pat} not found under {base}")
        return C.CDLL(m[0], mode=C.RTLD_GLOBAL)
    avutil = one("libavutil-*.so.*")
    try:
        one("libswresample-*.so.*")
    except ImportError:
        pass
    avcodec = one("libavcodec-*.so.*")
    avformat = one("libavformat-*.so.*")
    vp = C.c_void_p
    avformat.avformat_open_input.argtypes = [C.POINTER(vp), C.c_char_p, vp, vp]
    avformat.avformat_find_stream_info.argtypes = [vp, vp]
    avformat.av_read_frame.argtypes = [vp, vp]
    avformat.avformat_close_input.argtypes = [C.POINTER(vp)]
    avcodec.avcodec_find_decoder_by_name.restype = vp
    avcodec.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    avcodec.avcodec_alloc_context3.restype = vp
    avcodec.avcodec_alloc_context3.argtypes = [vp]
    avcodec.avcodec_parameters_to_context.argtypes = [vp, vp]
    avcodec.avcodec_open2.argtypes = [vp, vp, vp]
    avcodec.avcodec_send_packet.argtypes = [vp, vp]
    avcodec.avcodec_receive_frame.argtypes = [vp, vp]
    avcodec.av_packet_alloc.restype = vp
    avcodec.av_packet_unref.argtypes = [vp]
    avcodec.av_packet_free.argtypes = [C.POINTER(vp)]
    avcodec.avcodec_free_context.argtypes = [C.POINTER(vp)]
    avutil.av_frame_alloc.restype = vp
    avutil.av_frame_free.argtypes = [C.POINTER(vp)]
    avutil.av_frame_unref.argtypes = [vp]
    _libs = (avutil, avcodec, avformat)
    return _libs


def _rd(addr, ctype):
    return ctype.from_address(addr).value


def decode_file(path: str, channels: int) -> np.ndarray:
    """Decode a FLAC file -> (N, channels) int32 (s16 output is widened; s32 is shifted down when
    FFmpeg left-justifies <32-bit samples)."""
    avutil, avcodec, avformat = _load()
    fmt = C.c_void_p(None)
    if avformat.avformat_open_input(C.byref(fmt), path.encode(), None, None) < 0:
        raise ValueError("avformat_open_input failed")
    try:
        if avformat.avformat_find_stream_info(fmt, None) < 0:
            raise ValueError("find_stream_info failed")
        streams = _rd(fmt.value + 48, C.c_void_p)
        st0 = _rd(streams, C.c_void_p)
        codecpar = _rd(st0 + 16, C.c_void_p)
        dec = avcodec.avcodec_find_decoder_by_name(b"flac")
        ctx = C.c_void_p(avcodec.avcodec_alloc_context3(dec))
        avcodec.avcodec_parameters_to_context(ctx, codecpar)
        if avcodec.avcodec_open2(ctx, dec, None) < 0:
            raise ValueError("avcodec_open2 failed")
        pkt = C.c_void_p(avcodec.av_packet_alloc())
        frm = C.c_void_p(avutil.av_frame_alloc())
        chunks = []

        def drain():
            while avcodec.avcodec_receive_frame(ctx, frm) == 0:
                n = _rd(frm.value + 112, C.c_int)
                f = _rd(frm.value + 116, C.c_int)
                ext = _rd(frm.value + 96, C.c_void_p)
                p0 = _rd(ext, C.c_void_p)
                if f == 1:      # AV_SAMPLE_FMT_S16 packed
                    a = np.ctypeslib.as_array((C.c_int16 * (n * channels)).from_address(p0)).astype(np.int32)
                    chunks.append(a.reshape(n, channels).copy())
                elif f == 2:    # AV_SAMPLE_FMT_S32 packed
                    a = np.ctypeslib.as_array((C.c_int32 * (n * channels)).from_address(p0))
                    chunks.append(a.reshape(n, channels).copy())
                elif f in (6, 7):   # planar s16 / s32
                    ct, dt = (C.c_int16, np.int16) if f == 6 else (C.c_int32, np.int32)
                    cols = []
                    for c in range(channels):
                        pc = _rd(ext + 8 * c, C.c_void_p)
                        cols.append(np.ctypeslib.as_array((ct * n).from_address(pc)).astype(np.int32))
                    chunks.append(np.stack(cols, axis=1))
                else:
                    raise ValueError(f"unexpected sample format {f}")
                avutil.av_frame_unref(frm)

        while avformat.av_read_frame(fmt, pkt) >= 0:
            if avcodec.avcodec_send_packet(ctx, pkt) < 0:
                avcodec.av_packet_unref(pkt)
                raise ValueError("send_packet failed (corrupt frame?)")
            avcodec.av_packet_unref(pkt)
            drain()
        avcodec.avcodec_send_packet(ctx, None)
        drain()
        avutil.av_frame_free(C.byref(frm))
        avcodec.av_packet_free(C.byref(pkt))
        avcodec.avcodec_free_context(C.byref(ctx))
    finally:
        avformat.avformat_close_input(C.byref(fmt))
    if not chunks:
        return np.zeros((0, channels), dtype=np.int32)
    return np.concatenate(chunks, axis=0)


def decode_bytes(data: bytes, channels: int) -> np.ndarray:
    with tempfile.NamedTemporaryFile(suffix=".flac", delete=False) as tmp:
        tmp.write(data)
        name = tmp.name
    try:
        return decode_file(name, channels)
    finally:
        os.unlink(name)


def encode_timing(path: str, level: int = 5, blocksize: int = 4096, repeat: int = 1):
    """CPU-baseline helper (bench.py cpu_baseline.ffmpeg): decode `path` with libavcodec, keep the decoded AVFrames, then
    TIME libavcodec's FLAC ENCODER (SIMD LPC / residual code, the only production FLAC encoder on the box) over them at
    `level`.  The frames the decoder produced carry format, channel layout, rate and sample count, so no AVFrame field is
    written by offset; the encoder context is filled from the file's codec parameters and three AVOptions.
    Returns (seconds, samples_per_channel, encoded_bytes)."""
    import time
    avutil, avcodec, avformat = _load()
    vp = C.c_void_p
    avcodec.avcodec_find_encoder_by_name.restype = vp
    avcodec.avcodec_find_encoder_by_name.argtypes = [C.c_char_p]
    avcodec.avcodec_send_frame.argtypes = [vp, vp]
    avcodec.avcodec_receive_packet.argtypes = [vp, vp]
    avutil.av_frame_clone.restype = vp
    avutil.av_frame_clone.argtypes = [vp]
    avutil.av_opt_set_int.argtypes = [vp, C.c_char_p, C.c_int64, C.c_int]
    avutil.av_opt_set.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
    fmt = vp(None)
    if avformat.avformat_open_input(C.byref(fmt), path.encode(), None, None) < 0:
        raise ValueError("avformat_open_input failed")
    frames = []
    try:
        if avformat.avformat_find_stream_info(fmt, None) < 0:
            raise ValueError("find_stream_info failed")
        streams = _rd(fmt.value + 48, vp)
        st0 = _rd(streams, vp)
        codecpar = _rd(st0 + 16, vp)
        dec = avcodec.avcodec_find_decoder_by_name(b"flac")
        dctx = vp(avcodec.avcodec_alloc_context3(dec))
        avcodec.avcodec_parameters_to_context(dctx, codecpar)
        if avcodec.avcodec_open2(dctx, dec, None) < 0:
            raise ValueError("avcodec_open2(decoder) failed")
        pkt = vp(avcodec.av_packet_alloc())
        frm = vp(avutil.av_frame_alloc())
        n_samples = 0

        def drain():
            nonlocal n_samples
            while avcodec.avcodec_receive_frame(dctx, frm) == 0:
                n_samples += _rd(frm.value + 112, C.c_int)
                frames.append(vp(avutil.av_frame_clone(frm)))
                avutil.av_frame_unref(frm)

        while avformat.av_read_frame(fmt, pkt) >= 0:
            rc = avcodec.avcodec_send_packet(dctx, pkt)
            avcodec.av_packet_unref(pkt)
            if rc < 0:
                raise ValueError("send_packet failed")
            drain()
        avcodec.avcodec_send_packet(dctx, None)
        drain()
        enc = avcodec.avcodec_find_encoder_by_name(b"flac")
        if not enc:
            raise ValueError("this libavcodec has no FLAC encoder")
        total_s, out_bytes = 0.0, 0
        for _ in range(repeat):
            ectx = vp(avcodec.avcodec_alloc_context3(enc))
            avcodec.avcodec_parameters_to_context(ectx, codecpar)        # rate, channel layout, sample format, bits per raw sample
            avutil.av_opt_set_int(ectx, b"compression_level", level, 0)
            avutil.av_opt_set_int(ectx, b"frame_size", blocksize, 0)
            avutil.av_opt_set(ectx, b"time_base", b"1/%d" % 44100, 0)
            if avcodec.avcodec_open2(ectx, enc, None) < 0:
                raise ValueError("avcodec_open2(encoder) failed")
            out_bytes = 0
            t0 = time.perf_counter()
            for f in frames:
                if avcodec.avcodec_send_frame(ectx, f) < 0:
                    raise ValueError("send_frame failed")
                while avcodec.avcodec_receive_packet(ectx, pkt) == 0:
                    out_bytes += _rd(pkt.value + 32, C.c_int)
                    avcodec.av_packet_unref(pkt)
            avcodec.avcodec_send_frame(ectx, None)
            while avcodec.avcodec_receive_packet(ectx, pkt) == 0:
                out_bytes += _rd(pkt.value + 32, C.c_int)
                avcodec.av_packet_unref(pkt)
            total_s += time.perf_counter() - t0
            avcodec.avcodec_free_context(C.byref(ectx))
        avutil.av_frame_free(C.byref(frm))
        avcodec.av_packet_free(C.byref(pkt))
        avcodec.avcodec_free_context(C.byref(dctx))
    finally:
        for f in frames:
            avutil.av_frame_free(C.byref(f))
        avformat.avformat_close_input(C.byref(fmt))
    return total_s / repeat, n_samples, out_bytes
