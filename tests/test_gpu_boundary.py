"""GPU suite: the libFLAC-shaped handle API of include/flacraster_b200.h section 6, driven through ctypes exactly as
pyflac's cffi layer drives FLAC__stream_encoder_* / FLAC__stream_decoder_* (docs/sonos-pyflac.txt:2186-2212, :1994-2014,
:1584-1629, :1809-1854).  Encoder output must be libFLAC's golden bytes; decoder output the oracle's samples."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    from flac_raster_b200 import _native
    _native.build()
    _native.require_cuda()
    return _native


def _encode_with_handle(nat, pcm, bps, rate, level, seekable=False, verify=False, planar=False):
    """Returns (chunks [(bytes, samples, current_frame)], finish rc, final state, metadata seen)."""
    L = nat.lib()
    out = bytearray()
    pos = [0]
    calls = []
    meta = []

    def wcb(enc, buf, nbytes, samples, cur, client):
        data = C.string_at(buf, nbytes)
        calls.append((data, samples, cur))
        end = pos[0] + nbytes
        if end > len(out):
            out.extend(b"\0" * (end - len(out)))
        out[pos[0]:end] = data
        pos[0] = end
        return 0

    def scb(enc, off, client):
        pos[0] = off
        return 0

    def tcb(enc, poff, client):
        poff[0] = pos[0]
        return 0

    def mcb(enc, md, client):
        si = md.contents.stream_info
        meta.append((md.contents.type, si.sample_rate, si.channels, si.bits_per_sample, si.total_samples, si.min_framesize, si.max_framesize))

    w, s_, t_, m_ = nat.WRITE_CB(wcb), nat.ENC_SEEK_CB(scb), nat.ENC_TELL_CB(tcb), nat.ENC_METADATA_CB(mcb)
    e = L.frb_stream_encoder_new()
    assert e
    try:
        assert L.frb_stream_encoder_get_state(e) == 1                       # UNINITIALIZED
        assert L.frb_stream_encoder_set_verify(e, 1 if verify else 0) == 1
        assert L.frb_stream_encoder_set_channels(e, pcm.shape[1]) == 1
        assert L.frb_stream_encoder_set_bits_per_sample(e, bps) == 1
        assert L.frb_stream_encoder_set_sample_rate(e, rate) == 1
        assert L.frb_stream_encoder_set_compression_level(e, level) == 1
        assert L.frb_stream_encoder_set_blocksize(e, 4096) == 1
        assert L.frb_stream_encoder_set_streamable_subset(e, 1) == 1
        assert L.frb_stream_encoder_set_limit_min_bitrate(e, 0) == 1
        P = nat.cb_ptr
        rc = L.frb_stream_encoder_init_stream(e, P(w), P(s_) if seekable else None, P(t_) if seekable else None, P(m_) if seekable else None, None)
        assert rc == 0 and L.frb_stream_encoder_get_state(e) == 0
        assert L.frb_stream_encoder_set_channels(e, 2) == 0               # setters are refused once initialised (libFLAC)
        x = np.ascontiguousarray(pcm, dtype=np.int32)
        half = (x.shape[0] // 2 // 4096) * 4096 + 17                        # two calls, split off a block boundary
        if planar:
            for a, b in ((0, half), (half, x.shape[0])):
                cols = [np.ascontiguousarray(x[a:b, c]) for c in range(x.shape[1])]
                ptrs = (C.POINTER(C.c_int32) * len(cols))(*[c.ctypes.data_as(C.POINTER(C.c_int32)) for c in cols])
                assert L.frb_stream_encoder_process(e, ptrs, b - a) == 1
        else:
            for a, b in ((0, half), (half, x.shape[0])):
                part = np.ascontiguousarray(x[a:b])
                assert L.frb_stream_encoder_process_interleaved(e, part.ctypes.data, b - a) == 1
        ok = L.frb_stream_encoder_finish(e)
        return bytes(out), calls, ok, L.frb_stream_encoder_get_state(e), meta
    finally:
        L.frb_stream_encoder_delete(e)


def test_encoder_handle_reproduces_libflac_golden_bytes(nat, rgb_pcm):
    golden = (GOLDEN / "sample_rgb.flac").read_bytes()
    blob, calls, ok, state, _ = _encode_with_handle(nat, rgb_pcm, 16, 44100, 5)
    assert ok == 1 and state == 1
    # callback contract (stream_encoder.h text, docs/sonos-pyflac.txt:6601-6634): header chunks first with samples == 0,
    # then one call per frame with samples == blocksize and current_frame == frame index
    heads = [c for c in calls if c[1] == 0]
    frames = [c for c in calls if c[1] != 0]
    assert calls[:len(heads)] == heads and heads[0][0] == b"fLaC" and len(heads[1][0]) == 38
    assert [c[2] for c in frames] == list(range(16)) and all(c[1] == 4096 for c in frames)
    assert b"".join(c[0] for c in frames) == golden[86:]                     # libFLAC 1.4.3's own 178 857 frame bytes
    # same through the planar entry point
    blob2, calls2, ok2, _, _ = _encode_with_handle(nat, rgb_pcm, 16, 44100, 5, planar=True)
    assert ok2 == 1 and blob2 == blob


def test_encoder_handle_seekable_finalises_streaminfo_and_verifies(nat, oracle, rgb_pcm):
    from flac_raster_b200 import flacfmt
    blob, calls, ok, state, meta = _encode_with_handle(nat, rgb_pcm, 16, 44100, 5, seekable=True, verify=True)
    assert ok == 1 and state == 1
    si = flacfmt.parse_header(blob).streaminfo
    sizes = [len(c[0]) for c in calls if c[1] != 0]
    assert (si.total_samples, si.min_framesize, si.max_framesize) == (65536, min(sizes), max(sizes))
    assert meta == [(0, 44100, 3, 16, 65536, min(sizes), max(sizes))]
    dec, info = oracle.decode(blob)                                          # still a valid stream, now with its length
    assert np.array_equal(dec, rgb_pcm)
    plain, _, _, _, _ = _encode_with_handle(nat, rgb_pcm, 16, 44100, 5)
    assert blob[:8] == plain[:8] and blob[42:] == plain[42:]                 # only the STREAMINFO body was rewritten


def test_encoder_handle_rejects_bad_settings(nat):
    L = nat.lib()
    w = nat.WRITE_CB(lambda *a: 0)
    for setter, value, want in (("channels", 9, 4), ("bits_per_sample", 24, 5), ("sample_rate", 0, 6), ("blocksize", 8192, 7),
                                ("limit_min_bitrate", 1, 1)):
        e = L.frb_stream_encoder_new()
        getattr(L, f"frb_stream_encoder_set_{setter}")(e, value)
        assert L.frb_stream_encoder_init_stream(e, nat.cb_ptr(w), None, None, None, None) == want, setter
        assert L.frb_stream_encoder_process_interleaved(e, None, 0) == 0     # not initialised
        L.frb_stream_encoder_delete(e)
    e = L.frb_stream_encoder_new()
    assert L.frb_stream_encoder_init_stream(e, None, None, None, None, None) == 3    # INVALID_CALLBACKS
    assert L.frb_stream_encoder_init_stream(e, nat.cb_ptr(w), None, None, None, None) == 0
    assert L.frb_stream_encoder_init_stream(e, nat.cb_ptr(w), None, None, None, None) == 13      # ALREADY_INITIALIZED
    assert L.frb_stream_encoder_finish(e) == 1                                        # no samples: header only
    L.frb_stream_encoder_delete(e)


def _decode_with_handle(nat, path=None, data=None, abort_after=None):
    L = nat.lib()
    frames, metas, errors = [], [], []

    def wcb(dec, frame, buffer, client):
        h = frame.contents.header
        chans = [np.ctypeslib.as_array(buffer[c], shape=(h.blocksize,)).copy() for c in range(h.channels)]
        frames.append((h.blocksize, h.sample_rate, h.channels, h.bits_per_sample, h.number.frame_number, np.column_stack(chans)))
        return 1 if abort_after is not None and len(frames) >= abort_after else 0

    def mcb(dec, md, client):
        si = md.contents.stream_info
        metas.append((md.contents.type, si.sample_rate, si.channels, si.bits_per_sample, si.max_blocksize, si.total_samples))

    def ecb(dec, status, client):
        errors.append(status)

    pos = [0]

    def rcb(dec, buf, pbytes, client):
        n = min(pbytes[0], len(data) - pos[0], 70001)                     # odd chunk size on purpose
        C.memmove(buf, data[pos[0]:pos[0] + n], n)
        pos[0] += n
        pbytes[0] = n
        return 1 if pos[0] >= len(data) else 0

    w, m, er, r = nat.DEC_WRITE_CB(wcb), nat.DEC_METADATA_CB(mcb), nat.DEC_ERROR_CB(ecb), nat.DEC_READ_CB(rcb)
    d = L.frb_stream_decoder_new()
    try:
        assert L.frb_stream_decoder_get_state(d) == 9                         # UNINITIALIZED
        if path is not None:
            rc = L.frb_stream_decoder_init_file(d, str(path).encode(), nat.cb_ptr(w), nat.cb_ptr(m), nat.cb_ptr(er), None)
        else:
            rc = L.frb_stream_decoder_init_stream(d, nat.cb_ptr(r), None, None, None, None, nat.cb_ptr(w), nat.cb_ptr(m), nat.cb_ptr(er), None)
        if rc != 0:
            return rc, None, None, frames, metas, errors, None
        ok = L.frb_stream_decoder_process_until_end_of_stream(d)
        state = L.frb_stream_decoder_get_state(d)
        info = (L.frb_stream_decoder_get_channels(d), L.frb_stream_decoder_get_bits_per_sample(d), L.frb_stream_decoder_get_sample_rate(d),
                L.frb_stream_decoder_get_blocksize(d), L.frb_stream_decoder_get_total_samples(d))
        assert L.frb_stream_decoder_finish(d) == 1 and L.frb_stream_decoder_get_state(d) == 9
        return rc, ok, state, frames, metas, errors, info
    finally:
        L.frb_stream_decoder_delete(d)


def test_decoder_handle_file_and_stream_on_reference_golden(nat, rgb_pcm):
    path = GOLDEN / "sample_rgb.flac"
    for kw in ({"path": path}, {"data": path.read_bytes()}):
        rc, ok, state, frames, metas, errors, info = _decode_with_handle(nat, **kw)
        assert rc == 0 and ok == 1 and state == 4 and errors == []             # END_OF_STREAM
        assert metas == [(0, 44100, 3, 16, 4096, 0)]                            # reference files: total_samples 0
        assert [f[:5] for f in frames] == [(4096, 44100, 3, 16, i) for i in range(16)]
        assert np.array_equal(np.concatenate([f[5] for f in frames]), rgb_pcm)
        assert info == (3, 16, 44100, 4096, 65536)


def test_decoder_handle_legacy_multistream_tail_abort_and_errors(nat, oracle, tmp_path):
    # legacy --spatial golden: four streams back to back, the handle yields the first (like a libFLAC decoder fed this file)
    rc, ok, state, frames, metas, errors, info = _decode_with_handle(nat, path=GOLDEN / "sample_dem.flac")
    assert rc == 0 and ok == 1 and state == 4 and metas[0][1:4] == (44100, 1, 32)
    got = np.concatenate([f[5] for f in frames])
    want, _ = oracle.decode((GOLDEN / "sample_dem.flac").read_bytes()[:10426])
    assert np.array_equal(got, want) and got.shape == (65536, 1)
    # short tail frame + write callback abort
    x = (3000 * np.sin(np.arange(3 * 4096 + 123) / 11.0)).astype(np.int32).reshape(-1, 1)
    enc, _ = oracle.encode(x, 16, 48000, 5)
    rc, ok, state, frames, _, _, info = _decode_with_handle(nat, data=bytes(enc))
    assert ok == 1 and [f[0] for f in frames] == [4096, 4096, 4096, 123] and np.array_equal(np.concatenate([f[5] for f in frames]), x)
    rc, ok, state, frames, _, _, _ = _decode_with_handle(nat, data=bytes(enc), abort_after=2)
    assert ok == 0 and state == 7 and len(frames) == 2                          # ABORTED
    # a damaged frame is reported through the error callback (FRAME_CRC_MISMATCH = 2), nothing is written
    bad = bytearray(enc)
    bad[len(bad) // 2] ^= 0x10
    rc, ok, state, frames, _, errors, _ = _decode_with_handle(nat, data=bytes(bad))
    assert ok == 0 and state == 7 and frames == [] and errors and errors[0] in (0, 2)
    # missing file / not a FLAC stream
    assert _decode_with_handle(nat, path=tmp_path / "nope.flac")[0] == 4     # ERROR_OPENING_FILE
    rc, ok, state, frames, metas, errors, _ = _decode_with_handle(nat, data=b"RIFF" + bytes(100))
    assert frames == [] and metas == [] and errors == [0]
