"""CPU suite part 2: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import json
import re
import struct
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def test_library_loads_and_exports_every_declared_symbol():
    """Every function declared in include/flacraster_b200.h is exported by the built .so."""
    from flac_raster_b200 import _native as nat
    nat.build()
    hdr = (ROOT / "include" / "flacraster_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(frb_[a-z0-9_]+)\s*\(", hdr))
    declared = {d for d in declared if not d.endswith("_cb")}          # callback typedefs
    assert len(declared) >= 30
    lib = ctypes.CDLL(str(nat.LIB_PATH))
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(nat._EXPORTS) <= declared
    assert nat.lib().frb_version() >= 1
    assert nat.lib().frb_error_string(5) == b"CRC mismatch"


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: compute entry points raise when no CUDA device is present."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from flac_raster_b200 import _native as nat, normalize_to_audio
    with pytest.raises(nat.NativeError):
        nat.host_encode(np.zeros((10, 1), np.int32), 16, 44100)
    with pytest.raises(nat.NativeError):
        normalize_to_audio(np.arange(10, dtype=np.uint8), 16)


def test_product_never_imports_oracle():
    for p in (ROOT / "flac_raster_b200").rglob("*.py"):
        src = p.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), p


def test_flac_header_parse_golden_files():
    from flac_raster_b200 import flacfmt
    h = flacfmt.parse_header((GOLDEN / "sample_rgb.flac").read_bytes())
    si = h.streaminfo
    assert (si.sample_rate, si.channels, si.bits_per_sample, si.max_blocksize, si.total_samples) == (44100, 3, 16, 4096, 0)
    assert h.vendor == "reference libFLAC 1.4.3 20230623" and h.first_frame_offset == 86
    h = flacfmt.parse_header((GOLDEN / "sample_dem.flac").read_bytes())
    assert h.streaminfo.bits_per_sample == 32 and h.first_frame_offset == 2058
    assert h.tags["GEOSPATIAL_CRS"] == ["EPSG:4326"]
    assert "GEOSPATIAL_SPATIAL_INDEX" in h.tags


def test_flac_header_roundtrip_and_oracle_accepts_it(oracle):
    from flac_raster_b200 import flacfmt
    si = flacfmt.StreamInfo(4096, 4096, 0, 0, 192000, 8, 32, (1 << 35) + 5)
    blob = flacfmt.build_header(si, {"GEOSPATIAL_CRS": "EPSG:4326", "K": "v=1"}, padding=100)
    h = flacfmt.parse_header(blob + b"\xff\xf8")
    assert h.streaminfo == si and h.tags == {"GEOSPATIAL_CRS": ["EPSG:4326"], "K": ["v=1"]}
    assert h.first_frame_offset == len(blob)
    # a header written by flacfmt followed by oracle frames is a stream the oracle decodes
    x = (1000 * np.sin(np.arange(9000) / 7.0)).astype(np.int32).reshape(-1, 1)
    enc, fs = oracle.encode(x, 16, 44100, 5)
    frames = enc[len(enc) - int(fs.sum()):]
    mine = flacfmt.build_header(flacfmt.StreamInfo(4096, 4096, 0, 0, 44100, 1, 16, 9000), {"A": "b"}) + frames
    dec, _ = oracle.decode(mine)
    assert np.array_equal(dec, x)


def test_metadata_tags_roundtrip():
    from flac_raster_b200.converter import metadata_tags, parse_metadata_tags, tile_metadata
    md = tile_metadata(740, 512, 3, "uint16", "EPSG:32633", (10.0, 0.0, 5e5, 0.0, -10.0, 4e6), 0.0, 11672.0, None, 32767)
    tags = {k: [v] for k, v in metadata_tags(md).items()}
    back = parse_metadata_tags(tags)
    assert back["width"] == 740 and back["height"] == 512 and back["count"] == 3 and back["dtype"] == "uint16"
    assert back["data_min"] == 0.0 and back["data_max"] == 11672.0 and back["nodata"] is None
    assert back["transform"][:6] == [10.0, 0.0, 5e5, 0.0, -10.0, 4e6] and len(back["transform"]) == 9
    assert back["bounds"] == {"left": 5e5, "bottom": 4e6 - 5120.0, "right": 5e5 + 7400.0, "top": 4e6}
    # float repr round-trips exactly
    md2 = tile_metadata(1, 1, 1, "float32", None, None, 0.1 + 0.2, -1e-300, -9999.0, 8388607)
    b2 = parse_metadata_tags({k: [v] for k, v in metadata_tags(md2).items()})
    assert b2["data_min"] == 0.1 + 0.2 and b2["data_max"] == -1e-300 and b2["nodata"] == -9999.0


def test_tiff_roundtrip(tmp_path):
    from flac_raster_b200.tiffio import read_geotiff, write_geotiff
    r = read_geotiff(GOLDEN / "sample_dem.tif")
    assert r.data.shape == (1, 512, 512) and r.data.dtype == np.int16 and r.crs == "EPSG:4326"
    assert r.transform == (0.001, 0.0, -105.5, 0.0, -0.001, 40.5)
    rng = np.random.default_rng(0)
    for dt in ("uint8", "int8", "uint16", "int16", "uint32", "int32", "float32", "float64"):
        a = (rng.random((3, 19, 23)) * 100).astype(dt)
        p = tmp_path / f"{dt}.tif"
        write_geotiff(p, a, (2.0, 0.0, 100.0, 0.0, -2.0, 50.0), "EPSG:32633", -1.0)
        b = read_geotiff(p)
        assert np.array_equal(a, b.data) and b.data.dtype == a.dtype
        assert b.transform == (2.0, 0.0, 100.0, 0.0, -2.0, 50.0) and b.crs == "EPSG:32633" and b.nodata == -1.0


def test_tile_grid_matches_reference_enumeration():
    from flac_raster_b200.engine import tile_grid
    t = tile_grid(10980, 10980, 1024)                     # config C3: 121 tiles, edge 740
    assert len(t) == 121
    assert (t[0]["h"], t[0]["w"]) == (1024, 1024) and (t[10]["h"], t[10]["w"]) == (1024, 740)
    assert (t[120]["row_off"], t[120]["col_off"], t[120]["h"], t[120]["w"]) == (10240, 10240, 740, 740)
    assert sum(int(a["h"]) * int(a["w"]) for a in t) == 10980 * 10980
    # row-major order, frame_id increments along columns first (cli.py:553-554)
    assert (t[1]["row_off"], t[1]["col_off"]) == (0, 1024) and (t[11]["row_off"], t[11]["col_off"]) == (1024, 0)


def test_audio_params_table():
    from flac_raster_b200.normalization import audio_params_for, estimate_precision_loss
    assert audio_params_for((512, 512), "int16") == (44100, 16)
    assert audio_params_for((1024, 1024), "uint16") == (48000, 16)
    assert audio_params_for((8, 10980, 10980), "uint16") == (192000, 16)
    assert audio_params_for((32768, 32768), "float32") == (192000, 24)
    assert audio_params_for((740, 1024), "uint8") == (44100, 16)
    e = estimate_precision_loss("uint16", 0.0, 65535.0, 16)
    assert e["quantization_levels"] == 65534 and not e["is_lossless"]
    assert estimate_precision_loss("uint8", 0, 255, 16)["is_lossless"]


def test_spatial_index_and_streamer_on_synthetic_container(tmp_path):
    """Index/byte-range logic of the streamer (no decode): strict bbox test, merged ranges."""
    from flac_raster_b200.spatial_encoder import SpatialFLACStreamer
    frames = []
    off = 0
    for i in range(4):
        r, c = divmod(i, 2)
        frames.append({"frame_id": i, "bbox": [c * 10.0, 20.0 - (r + 1) * 10.0, (c + 1) * 10.0, 20.0 - r * 10.0],
                       "window": {"col_off": c * 10, "row_off": r * 10, "width": 10, "height": 10},
                       "byte_offset": off, "byte_size": 100 + i})
        off += 100 + i
    index = {"crs": "EPSG:4326", "transform": [1, 0, 0, 0, -1, 20, 0, 0, 1], "width": 20, "height": 20, "bands": 1,
             "dtype": "uint8", "tile_size": 10, "frames": frames}
    js = json.dumps(index, separators=(",", ":")).encode()
    p = tmp_path / "c.flac"
    p.write_bytes(len(js).to_bytes(4, "big") + js + bytes(off))
    s = SpatialFLACStreamer(p)
    assert s.header_size == 4 + len(js) and len(s.spatial_index.frames) == 4
    assert [f.frame_id for f in s.spatial_index.query_bbox((0, 0, 20, 20))] == [0, 1, 2, 3]
    assert [f.frame_id for f in s.spatial_index.query_bbox((10, 10, 20, 20))] == [1]      # touching edges excluded
    assert s.get_byte_ranges_for_bbox((0, 10.5, 20, 20)) == [(s.header_size, s.header_size + 200)]
    assert s.get_byte_ranges_for_bbox((100, 100, 101, 101)) == []
    assert len(s.stream_bbox_data((0, 0, 5, 5))) == 102
    with pytest.raises(FileNotFoundError):
        SpatialFLACStreamer(tmp_path / "missing.flac")


def test_shard_range_partitions():
    from flac_raster_b200.distributed import exclusive_scan, shard_range
    for n, w in ((121, 8), (4096, 8), (3, 8), (1, 1), (484, 4)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert list(exclusive_scan(np.array([5, 7, 9]))) == [0, 5, 12]


def test_weighted_shards_balance_pixel_counts():
    """distributed.shard_ranges_weighted / tile_shards: contiguous blocks, every item once, the heaviest rank as light as a
    contiguous split allows (checked against a brute-force search on small cases), equal weights = equal counts."""
    import itertools
    from flac_raster_b200.distributed import shard_range, shard_ranges_weighted, tile_shards
    from flac_raster_b200.engine import tile_grid
    rng = np.random.default_rng(5)
    for n, world in ((1, 1), (1, 4), (3, 8), (7, 3), (9, 4), (10, 2), (12, 5)):
        w = rng.integers(1, 50, n).tolist()
        spans = shard_ranges_weighted(w, world)
        assert len(spans) == world and spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and all(b >= a for a, b in spans)
        assert sum(1 for a, b in spans if b > a) == min(n, world)                 # no idle rank while items are left
        k = min(n, world)
        best = min(max(sum(w[a:b]) for a, b in zip((0,) + cuts, cuts + (n,))) for cuts in itertools.combinations(range(1, n), k - 1))
        assert max(sum(w[a:b]) for a, b in spans) == best
    for n, world in ((4096, 8), (121, 8), (16, 4)):
        assert max(b - a for a, b in shard_ranges_weighted([262144] * n, world)) == -(-n // world)
    # C3: 121 tiles with a 740-pixel last column and row; N = 8: 2.4 % above the mean instead of 9.4 % by tile count
    tiles = tile_grid(10980, 10980, 1024)
    px = tiles["h"].astype(np.int64) * tiles["w"]
    for world, bound in ((2, 1.003), (4, 1.015), (8, 1.025)):
        spans = tile_shards(tiles, world)
        assert max(px[a:b].sum() for a, b in spans) <= bound * px.sum() / world
        assert max(px[a:b].sum() for a, b in (shard_range(121, r, world) for r in range(world))) > max(px[a:b].sum() for a, b in spans)


def test_pyflac_shim_signatures():
    """Constructor/method surface of pyflac 3.0.0 that the reference uses."""
    import inspect
    from flac_raster_b200 import codec
    sig = inspect.signature(codec.StreamEncoder.__init__)
    assert list(sig.parameters)[1:] == ["sample_rate", "write_callback", "seek_callback", "tell_callback", "metadata_callback",
                                        "compression_level", "blocksize", "streamable_subset", "verify", "limit_min_bitrate"]
    assert sig.parameters["compression_level"].default == 5 and sig.parameters["blocksize"].default == 0
    enc = codec.StreamEncoder(write_callback=lambda *a: None, sample_rate=44100, compression_level=5, blocksize=4096)
    enc._channels, enc._bits_per_sample = 3, 24                 # the reference's dead writes (converter.py:147-148)
    with pytest.raises(TypeError):
        enc.process([1, 2, 3])
    with pytest.raises(codec.DecoderInitException):
        codec.FileDecoder("/nonexistent/file.flac")
    assert enc.finish() is False


def test_sample_rates_vector_rule_equals_reference_rule():
    """Engine.encode_tiles derives every tile's sample rate with one vector expression; it must be the reference's
    threshold rule (normalization.py:113-120) at and around every boundary."""
    from flac_raster_b200.normalization import audio_params_for, sample_rates_for_pixel_counts
    counts = [1, 999_999, 1_000_000, 1_000_001, 9_999_999, 10_000_000, 99_999_999, 100_000_000, 3_000_000_000]
    got = sample_rates_for_pixel_counts(np.array(counts))
    assert got.dtype == np.uint32
    assert [int(v) for v in got] == [audio_params_for((1, c), "uint16")[0] for c in counts]


def test_numa_binding_is_best_effort():
    """bind_to_gpu_numa_node never raises: without NVML / sysfs information it returns None and leaves the affinity."""
    import os
    from flac_raster_b200.distributed import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    node = bind_to_gpu_numa_node(0)
    assert node is None or isinstance(node, int)
    if node is None:
        assert os.sched_getaffinity(0) == before


def test_legacy_index_of_reference_golden_loads_and_offsets_are_repaired():
    """SURVEY Q6 / 8(f)4 (index part, no decode): the reference's legacy golden stores its index as base64(gzip(json)) with
    no marker tag and with byte offsets that went stale when mutagen grew stream 0; the streamer repairs them."""
    from flac_raster_b200 import flacfmt
    from flac_raster_b200.spatial_encoder import SpatialFLACStreamer
    path = GOLDEN / "sample_dem.flac"
    blob = path.read_bytes()
    s = SpatialFLACStreamer(path)
    fr = s.spatial_index.frames
    assert [(f.frame_id, f.byte_offset, f.byte_size) for f in fr] == [(0, 0, 10426), (1, 10426, 8454), (2, 18880, 8454), (3, 27334, 8454)]
    assert sum(f.byte_size for f in fr) == len(blob) == s.spatial_index.total_bytes
    for f in fr:
        h = flacfmt.parse_header(blob[f.byte_offset:f.byte_offset + f.byte_size])
        md = s._legacy_tile_metadata(f, h.streaminfo)
        assert (md["width"], md["height"], md["count"], md["dtype"], md["data_min"], md["data_max"]) == (256, 256, 1, "int16", 577.0, 1493.0)
        assert md["transform"][2] == -105.5 + f.window.col_off * 0.001 and md["transform"][5] == 40.5 - f.window.row_off * 0.001
    assert s.get_byte_ranges_for_bbox((-200, -90, 200, 90)) == [(0, len(blob) - 1)]


def test_legacy_index_plain_json_and_delta_repair(tmp_path):
    """Plain-JSON index text is accepted as well, and when the marker count does not match the index (a tile stream that
    happens to be missing) the growth of stream 0 alone repairs the later offsets."""
    from flac_raster_b200 import flacfmt
    from flac_raster_b200.spatial_encoder import SpatialFLACStreamer
    si = flacfmt.StreamInfo(4096, 4096, 0, 0, 44100, 1, 16, 0)
    body = [bytes([0xFF, 0xF8, i]) * (40 + i) for i in range(3)]
    plain = [flacfmt.build_header(si) + b for b in body]                     # what was appended while offsets were recorded
    offs, frames = 0, []
    for i, p in enumerate(plain):
        frames.append({"frame_id": i, "bbox": [i, 0.0, i + 1.0, 1.0], "window": {"row_off": 0, "col_off": i, "height": 1, "width": 1},
                       "byte_offset": offs, "byte_size": len(p)})
        offs += len(p)
    index = json.dumps({"crs": "EPSG:4326", "transform": [1, 0, 0, 0, -1, 1, 0, 0, 1], "frames": frames})
    tagged0 = flacfmt.build_header(si, {"GEOSPATIAL_CRS": "EPSG:4326", "GEOSPATIAL_SPATIAL_INDEX": index}, padding=333) + body[0]
    f = tmp_path / "legacy.flac"
    f.write_bytes(tagged0 + plain[1] + plain[2])
    s = SpatialFLACStreamer(f)
    blob = f.read_bytes()
    got = [(fr.byte_offset, fr.byte_size) for fr in s.spatial_index.frames]
    assert got == [(0, len(tagged0)), (len(tagged0), len(plain[1])), (len(tagged0) + len(plain[1]), len(plain[2]))]
    assert all(blob[o:o + 4] == b"fLaC" for o, _ in got)
    # index lists only two of the three streams: shift by the growth of stream 0
    index2 = json.dumps({"crs": "EPSG:4326", "transform": [1, 0, 0, 0, -1, 1, 0, 0, 1], "frames": frames[:2]})
    tagged0 = flacfmt.build_header(si, {"GEOSPATIAL_CRS": "EPSG:4326", "GEOSPATIAL_SPATIAL_INDEX": index2}) + body[0]
    f.write_bytes(tagged0 + plain[1] + plain[2])
    s = SpatialFLACStreamer(f)
    assert [(fr.byte_offset, fr.byte_size) for fr in s.spatial_index.frames] == [(0, len(tagged0)), (len(tagged0), len(plain[1]))]


def test_shard_plan_covers_every_tile_once():
    from flac_raster_b200.distributed import rows_of_shard, shard_plan, shard_range
    for world in (1, 2, 3, 8, 130):
        seen = []
        for r in range(world):
            tiles, (a, b), (r0, r1) = shard_plan(10980, 10980, 1024, r, world)
            seen += list(range(a, b))
            if b > a:
                assert r0 == int(tiles[a]["row_off"]) and r1 == int(tiles[b - 1]["row_off"]) + int(tiles[b - 1]["h"])
                assert (r1 - r0) <= 1024 * (2 + (b - a) // 11)
            else:
                assert (r0, r1) == (0, 0)
        assert seen == list(range(121))
    assert [shard_range(121, r, 8) for r in (0, 7)] == [(0, 16), (106, 121)]
    from flac_raster_b200.distributed import tile_shards
    assert [shard_plan(10980, 10980, 1024, r, 8)[1] for r in range(8)] == tile_shards(shard_plan(10980, 10980, 1024, 0, 8)[0], 8)


def test_c_tile_header_parser_equals_python_parser():
    """frb_parse_tile_headers / frb_gather_seek_index (host C, no GPU) against flacfmt.parse_header + parse_metadata_tags on
    tile files as the container holds them, on the reference goldens, and on damaged input."""
    from flac_raster_b200 import _native as nat, flacfmt
    from flac_raster_b200.converter import _DTYPE_NAMES, metadata_tags, parse_metadata_tags, tile_metadata
    L = nat.lib()
    rng = np.random.default_rng(3)
    files, want_idx = [], []
    cases = [("uint16", 3, 512, 300, 0.0, 11672.0, None), ("int16", 1, 512, 512, -3050.0, 3049.0, -9999.0),
             ("float32", 1, 77, 5, 0.1 + 0.2, 1e-300, float("nan")), ("uint8", 8, 1024, 1024, 7.0, 7.0, 255.0),
             ("float64", 2, 9, 4097, -1.5e300, 2.5e300, None)]
    for k, (dt, bands, w, h, mn, mx, nod) in enumerate(cases):
        md = tile_metadata(w, h, bands, dt, "EPSG:32633", (10.0, 0.0, 5e5 + k, 0.0, -10.0, 4e6), mn, mx, nod, 32767)
        bps = 16 if dt in ("uint8", "int8", "uint16", "int16") else 32
        nf = (w * h + 4095) // 4096
        fb = rng.integers(20, 9000, nf).astype(np.uint32)
        sb = rng.integers(40, 70000, nf * bands).astype(np.uint32)
        sidx = flacfmt.pack_seek_index(bands, 4096, fb, sb) if k != 2 else None
        si = flacfmt.StreamInfo(4096, 4096, 0, 0, 48000, bands, bps, w * h)
        files.append(flacfmt.build_header(si, metadata_tags(md), padding=33 if k == 1 else 0, seek_index=sidx) + b"\xff\xf8" + bytes(50 + k))
        want_idx.append((fb, sb if bands > 1 else None) if sidx is not None else None)
    files.append((GOLDEN / "sample_rgb.flac").read_bytes())          # reference file: mutagen-free vendor header, no GEOSPATIAL tags
    files.append((GOLDEN / "sample_dem.flac").read_bytes()[:10426])  # reference legacy stream 0: tags + padding written by mutagen
    blob = b"".join(files)
    offs = np.cumsum([0] + [len(f) for f in files[:-1]]).astype(np.uint64)
    sizes = np.array([len(f) for f in files], dtype=np.uint64)
    buf = np.frombuffer(blob, dtype=np.uint8)
    hdr = np.zeros(len(files), dtype=nat.TILE_HEADER_DTYPE)
    assert L.frb_parse_tile_headers(buf.ctypes.data, offs.ctypes.data, sizes.ctypes.data, len(files), hdr.ctypes.data) == 0
    for f, rec in zip(files, hdr):
        h = flacfmt.parse_header(f)
        si = h.streaminfo
        assert (rec["first_frame_offset"], rec["sample_rate"], rec["channels"], rec["bps"], rec["max_blocksize"], rec["total_samples"]) == \
               (h.first_frame_offset, si.sample_rate, si.channels, si.bits_per_sample, si.max_blocksize, si.total_samples)
        md = parse_metadata_tags(h.tags)
        assert bool(rec["flags"] & 1) == (md is not None)
        if md and "dtype" in md and md["dtype"] in _DTYPE_NAMES:
            assert (rec["width"], rec["height"], rec["count"], _DTYPE_NAMES[rec["dtype"]]) == (md["width"], md["height"], md["count"], md["dtype"])
            assert rec["data_min"] == md["data_min"] and rec["data_max"] == md["data_max"]
            nod = md.get("nodata")
            if nod is None:
                assert not rec["flags"] & 2
            else:
                assert rec["flags"] & 2 and (rec["nodata"] == nod or (np.isnan(nod) and np.isnan(rec["nodata"])))
        assert bool(rec["flags"] & 4) == (flacfmt.SEEK_INDEX_ID in h.applications)
        if rec["flags"] & 4:
            assert f[rec["index_offset"]:rec["index_offset"] + rec["index_len"]] == h.applications[flacfmt.SEEK_INDEX_ID]
    assert hdr[5]["dtype"] == -1 and hdr[6]["dtype"] == 3 and hdr[6]["data_min"] == 577.0      # the goldens
    # seek index gather: tiles 0 and 3 (3 and 8 channels differ -> one at a time), then a tile without a block
    for t in (0, 1, 3, 4):
        bands = int(hdr[t]["channels"])
        nf = np.array([want_idx[t][0].size], dtype=np.uint32)
        fb = np.zeros(int(nf[0]), dtype=np.uint32)
        sb = np.zeros(int(nf[0]) * bands, dtype=np.uint32)
        rc = L.frb_gather_seek_index(buf.ctypes.data, offs[t:].ctypes.data, hdr[t:].ctypes.data, 1, bands, 4096, nf.ctypes.data,
                                     fb.ctypes.data, sb.ctypes.data if bands > 1 else None)
        assert rc == 0 and np.array_equal(fb, want_idx[t][0]) and (bands == 1 or np.array_equal(sb, want_idx[t][1]))
        assert L.frb_gather_seek_index(buf.ctypes.data, offs[t:].ctypes.data, hdr[t:].ctypes.data, 1, bands, 1024, nf.ctypes.data,
                                       fb.ctypes.data, sb.ctypes.data) == nat.ERR_BAD_STREAM     # wrong blocksize
    one = np.array([1], dtype=np.uint32)
    assert L.frb_gather_seek_index(buf.ctypes.data, offs[2:].ctypes.data, hdr[2:].ctypes.data, 1, 1, 4096, one.ctypes.data,
                                   np.zeros(1, np.uint32).ctypes.data, None) == nat.ERR_BAD_STREAM
    # damaged: not a FLAC file, truncated metadata
    bad = np.frombuffer(b"RIFF" + bytes(60) + files[0][:40], dtype=np.uint8)
    boffs = np.array([0, 64], dtype=np.uint64)
    bsz = np.array([64, 40], dtype=np.uint64)
    bh = np.zeros(2, dtype=nat.TILE_HEADER_DTYPE)
    assert L.frb_parse_tile_headers(bad.ctypes.data, boffs.ctypes.data, bsz.ctypes.data, 2, bh.ctypes.data) == nat.ERR_BAD_STREAM
    assert bh[0]["first_frame_offset"] == 0 and bh[1]["first_frame_offset"] == 0


def test_remote_range_reads_into_a_buffer(range_http_server):
    """remote.RemoteFile against a local HTTP server (reference remote.py:137-177): inclusive ranges, keep-alive reuse,
    a server that ignores Range, and the streamer's index load + merged byte ranges over a URL (no decode here)."""
    from flac_raster_b200.remote import RemoteFile
    from flac_raster_b200.spatial_encoder import SpatialFLACStreamer
    base, d, log = range_http_server
    rng = np.random.default_rng(1)
    blob = rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes()
    (d / "f.bin").write_bytes(blob)
    (d / "norange_f.bin").write_bytes(blob)
    for name in ("f.bin", "norange_f.bin"):
        rf = RemoteFile(f"{base}/{name}")
        assert rf.read_range(10, 19) == blob[10:20]
        buf = bytearray(70_000)
        assert rf.read_range_into(1234, 1234 + 69_999, buf) == 70_000 and bytes(buf) == blob[1234:71_234]
        assert rf.read_range_into(299_990, 299_999, memoryview(buf)[:10]) == 10 and bytes(buf[:10]) == blob[-10:]
    assert "bytes=1234-71233" in log
    # a container index over HTTP
    frames, off = [], 0
    for i in range(3):
        frames.append({"frame_id": i, "bbox": [i, 0.0, i + 1.0, 1.0], "window": {"col_off": i, "row_off": 0, "width": 1, "height": 1},
                       "byte_offset": off, "byte_size": 1000 + i})
        off += 1000 + i
    js = json.dumps({"crs": "EPSG:4326", "transform": [1, 0, 0, 0, -1, 1, 0, 0, 1], "width": 3, "height": 1, "bands": 1, "dtype": "uint8",
                     "tile_size": 1, "frames": frames}, separators=(",", ":")).encode()
    (d / "c.flac").write_bytes(len(js).to_bytes(4, "big") + js + bytes(off))
    s = SpatialFLACStreamer(f"{base}/c.flac")
    assert s.is_url and s.header_size == 4 + len(js) and len(s.spatial_index.frames) == 3
    assert s.get_byte_ranges_for_bbox((0.5, 0.1, 2.5, 0.9)) == [(s.header_size, s.header_size + off - 1)]
    assert len(s.stream_bbox_data((1.1, 0.1, 1.9, 0.9))) == 1001


def test_oversized_tile_is_rejected():
    """A tile of 2^32 pixels or more would wrap the kernels' 32-bit pixel index: the engine refuses it up front."""
    from flac_raster_b200 import engine
    engine._check_tile_sizes(np.array([1 << 20, (1 << 32) - 1], dtype=np.int64))
    with pytest.raises(ValueError, match="pixels"):
        engine._check_tile_sizes(np.array([70000 * 70000], dtype=np.int64))


def test_full_tiles_first_ordering():
    """Host side of the tile-major result layout (converter.decode_staged_tiles): full-size tiles of a multi-band batch are decoded
    first so that their planes can be re-laid out on the device; the ragged rest keeps its relative order behind them."""
    from flac_raster_b200.converter import full_tiles_first
    # 3 x 4 grid of a 333 x 420 raster at tile 128: interior 128 x 128, last column 36 wide, last row 77 high
    w = np.array([128, 128, 128, 36] * 3)
    h = np.array([128] * 8 + [77] * 4)
    order, nf = full_tiles_first(w, h, 3)
    assert nf == 6 and list(order[:6]) == [0, 1, 2, 4, 5, 6] and list(order[6:]) == [3, 7, 8, 9, 10, 11]
    assert sorted(order) == list(range(12))
    assert full_tiles_first(w, h, 1) == (None, 0)                      # single band: views of the band-major block already
    assert full_tiles_first(w[:1], h[:1], 3) == (None, 0)              # one tile: the block is the tile
    assert full_tiles_first(np.array([64, 64, 64]), np.array([64, 64, 64]), 8) == (None, 3)          # all full: nothing moves
    assert full_tiles_first(np.array([64, 64, 10]), np.array([64, 64, 64]), 8) == (None, 2)          # full ones already lead
    assert full_tiles_first(np.array([64, 10, 20]), np.array([64, 64, 64]), 8) == (None, 0)          # a single full tile: not worth it
    o, nf = full_tiles_first(np.array([10, 64, 64]), np.array([64, 64, 64]), 2)
    assert nf == 2 and list(o) == [1, 2, 0]


def test_ctypes_bindings_match_the_header_arity():
    """Every entry point whose ctypes argtypes are declared in _native.py takes exactly as many parameters as its prototype in
    include/flacraster_b200.h (a binding that drifts from the header would pass garbage on the stack without any error)."""
    from flac_raster_b200 import _native as nat
    hdr = (ROOT / "include" / "flacraster_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(frb_[a-z0-9_]+)\s*\(", hdr):
        name = m.group(1)
        if name.endswith("_cb") or hdr[max(0, m.start() - 2):m.start()].endswith("(*"):
            continue
        depth, i, args, cur = 1, m.end(), [], ""
        while depth and i < len(hdr):
            ch = hdr[i]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    break
            if ch == "," and depth == 1:
                args.append(cur); cur = ""
            else:
                cur += ch
            i += 1
        args.append(cur)
        args = [a.strip() for a in args if a.strip() and a.strip() != "void"]
        if hdr[i + 1:i + 3].lstrip().startswith(";"):
            protos[name] = len(args)
    L = nat.lib()
    checked = 0
    for name, n_params in protos.items():
        fn = getattr(L, name, None)
        at = getattr(fn, "argtypes", None)
        if at is None:
            continue
        assert len(at) == n_params, (name, len(at), n_params)
        checked += 1
    assert checked >= 25, checked


def test_table_free_crc16_equals_bitwise_crc(oracle):
    """frb_selftest_crc16 runs the kernels' per-lane CRC-16 code (rows of 30 chunks folded with x^3840 = x^256 + 1 modulo the
    degree-15 factor of the polynomial, parity for the factor x + 1, CRT at the end; frb_crc16.cuh) on the host.  Checked against
    a bit-by-bit CRC-16/0x8005 and against the footer libFLAC wrote into the reference golden's frames, for every alignment of
    the start, lengths around the row size (480 bytes) and its multiples, and long ranges."""
    import ctypes as C
    from flac_raster_b200 import _native as nat
    L = nat.lib()

    def crc_ref(bs):
        c = 0
        for b in bs:
            c ^= b << 8
            for _ in range(8):
                c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
        return c

    rng = np.random.default_rng(11)
    raw = np.zeros(70000 + 64, dtype=np.uint8)
    base = (-raw.ctypes.data) % 16
    buf = raw[base:base + 70000 + 32]
    buf[:] = rng.integers(0, 256, buf.size, dtype=np.uint8)

    def crc_lib(a, e):
        out32, out128 = C.c_uint32(0), C.c_uint32(0)
        assert L.frb_selftest_crc16(buf.ctypes.data, a, e, 32, C.byref(out32)) == 0
        assert L.frb_selftest_crc16(buf.ctypes.data, a, e, 128, C.byref(out128)) == 0
        assert out32.value == out128.value, (a, e, out32.value, out128.value)
        return out32.value

    lengths = sorted(set(list(range(0, 40)) + [463, 464, 465, 479, 480, 481, 495, 496, 497, 511, 512, 513, 959, 960, 961, 976,
                                                  1440, 1919, 1920, 1921, 1936, 3839, 3840, 3841, 5760, 14400, 14401, 48211, 65535]))
    for a in list(range(0, 17)) + [31, 100, 1000]:
        for n in lengths:
            assert crc_lib(a, a + n) == crc_ref(buf[a:a + n].tobytes()), (a, n)
    # all-zero and all-ones payloads (parity / CRT corner cases)
    for fill in (0x00, 0xFF, 0x80, 0x01):
        buf[:20000] = fill
        for a, n in ((0, 480), (3, 481), (0, 14400), (5, 19000)):
            assert crc_lib(a, a + n) == crc_ref(buf[a:a + n].tobytes()), (fill, a, n)
    # libFLAC's own frames: the footers of the reference golden (frames found by walking 0xFFF8 candidates with the oracle's CRC)
    from flac_raster_b200 import flacfmt
    flac_oracle = oracle
    blob = (GOLDEN / "sample_rgb.flac").read_bytes()
    p0 = flacfmt.parse_header(blob).first_frame_offset
    fr = np.frombuffer(blob, dtype=np.uint8)
    frames = 0
    while p0 < len(blob):
        ends = [m.start() for m in re.finditer(b"\xff\xf8", blob[p0 + 2:])]
        ends = [p0 + 2 + q for q in ends] + [len(blob)]
        p1 = next(q for q in ends if flac_oracle.crc16(blob[p0:q - 2]) == ((blob[q - 2] << 8) | blob[q - 1]))
        for a in (0, 7):
            buf[a:a + p1 - p0] = fr[p0:p1]
            assert crc_lib(a, a + p1 - p0 - 2) == ((blob[p1 - 2] << 8) | blob[p1 - 1])
        frames += 1
        p0 = p1
    assert frames == 16


def test_size_exchange_unpack_with_pixel_count_shards():
    """The gathered size block of the sharded encode step (distributed.SizeExchange: `per` entries per rank, ranks hold
    tile_shards blocks of different lengths) unpacks to the sizes in tile order for the C3 grid at N = 2, 4, 8."""
    from flac_raster_b200.distributed import SizeExchange, exclusive_scan, tile_shards
    from flac_raster_b200.engine import tile_grid
    tiles = tile_grid(10980, 10980, 1024)
    sizes = np.arange(1000, 1000 + len(tiles), dtype=np.int64) * 7
    for world in (2, 4, 8):
        shards = tile_shards(tiles, world)
        x = SizeExchange(len(tiles), 0, world, "cpu", ranges=shards)
        assert x.per == max(b - a for a, b in shards) and x.recv_count == x.per * world
        recv = np.zeros((world, x.per), dtype=np.int64)
        for r, (a, b) in enumerate(shards):
            recv[r, :b - a] = sizes[a:b]
        got = x.unpack(recv.reshape(-1))
        assert np.array_equal(got, sizes)
        off = exclusive_scan(got)
        assert off[0] == 0 and off[-1] + got[-1] == sizes.sum()
