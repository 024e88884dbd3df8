"""Developer GPU check (not collected by pytest): quick parity sweep with verbose diagnostics.

    gpurun -- python tests/dev_gpu_check.py
"""
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from flac_raster_b200 import _native as nat          # noqa: E402
from flac_raster_b200 import flacfmt                 # noqa: E402
from oracle import flac_oracle as fo                 # noqa: E402

G = ROOT / "tests" / "golden"
results = []


def report(name, ok, extra=""):
    results.append((name, ok))
    print(("PASS " if ok else "FAIL ") + name + (" :: " + extra if extra else ""), flush=True)


def full_stream(payload, n, ch, bps, rate, blocksize=4096):
    si = flacfmt.StreamInfo(blocksize, blocksize, 0, 0, rate, ch, bps, n)
    return flacfmt.build_header(si) + bytes(payload)


def check_decode_golden():
    data = (G / "sample_rgb.flac").read_bytes()
    hdr = flacfmt.parse_header(data)
    ref, info = fo.decode(data)
    si = hdr.streaminfo
    for hint in (65536, 0):
        out = nat.host_decode(data[hdr.first_frame_offset:], si.channels, si.bits_per_sample, si.max_blocksize, si.sample_rate, hint)
        report(f"decode sample_rgb.flac hint={hint}", np.array_equal(out, ref), f"shape {out.shape}")
    data = (G / "sample_dem.flac").read_bytes()
    pos = 0
    k = 0
    while pos < len(data):
        ref, info = fo.decode(data[pos:])
        hdr = flacfmt.parse_header(data[pos:])
        si = hdr.streaminfo
        seg = data[pos + hdr.first_frame_offset: pos + info.bytes_consumed]
        out = nat.host_decode(seg, si.channels, si.bits_per_sample, si.max_blocksize, si.sample_rate, 0)
        report(f"decode sample_dem.flac stream {k} (32 bps)", np.array_equal(out, ref), f"shape {out.shape}")
        pos += info.bytes_consumed
        k += 1


def gen_cases():
    rng = np.random.default_rng(7)
    t = np.arange(50000)
    cases = {}
    cases["sine16_1ch"] = ((8000 * np.sin(t / 37.0) + 500 * rng.standard_normal(t.size)).astype(np.int32).reshape(-1, 1), 16)
    cases["noise16_3ch"] = (rng.integers(-32768, 32767, size=(20000, 3)).astype(np.int32), 16)
    cases["smooth16_8ch"] = ((np.cumsum(rng.integers(-20, 21, size=(30000, 8)), axis=0)).astype(np.int32), 16)
    cases["const16"] = (np.full((9000, 1), -32767, dtype=np.int32), 16)
    cases["wasted16"] = (((rng.integers(-2000, 2000, size=(12000, 2))) * 8).astype(np.int32), 16)
    cases["dem24_1ch"] = ((4e6 * np.sin(t / 300.0) + 2000 * rng.standard_normal(t.size)).astype(np.int32).reshape(-1, 1), 32)
    cases["noise24_2ch"] = (rng.integers(-8388607, 8388607, size=(10000, 2)).astype(np.int32), 32)
    cases["tiny3"] = (np.array([[1], [2], [-3]], dtype=np.int32), 16)
    cases["tail17"] = ((1000 * np.sin(np.arange(4096 + 17) / 9.0)).astype(np.int32).reshape(-1, 1), 16)
    cases["exact4096"] = ((1000 * np.sin(np.arange(8192) / 9.0)).astype(np.int32).reshape(-1, 1), 16)
    return cases


def check_encode():
    pcm, _ = fo.decode((G / "sample_rgb.flac").read_bytes())
    cases = gen_cases()
    cases["golden_rgb"] = (pcm, 16)
    for name, (x, bps) in cases.items():
        for level in (0, 2, 3, 5, 8):
            try:
                rate = 44100
                t0 = time.time()
                payload, fs = nat.host_encode(x, bps, rate, level)
                dt = time.time() - t0
                stream = full_stream(payload, x.shape[0], x.shape[1], bps, rate)
                dec, info = fo.decode(stream)
                ok = np.array_equal(dec, x)
                oenc, ofs = fo.encode(x, bps, rate, level, mid_side=(bps == 16))
                osz = int(ofs.sum())
                ratio = len(payload) / max(osz, 1)
                same = bytes(payload) == oenc[len(oenc) - osz:]
                report(f"encode {name} L{level}", ok and ratio <= 1.01 and int(fs.sum()) == len(payload),
                       f"bytes {len(payload)} oracle {osz} ratio {ratio:.4f} identical={same} {dt*1e3:.1f} ms")
                if ok:
                    back = nat.host_decode(payload, x.shape[1], bps, 4096, rate, x.shape[0])
                    report(f"  gpu-decode of gpu-encode {name} L{level}", np.array_equal(back, x))
            except Exception as e:  # noqa: BLE001
                traceback.print_exc()
                report(f"encode {name} L{level}", False, repr(e))


def check_oracle_streams_decode():
    """GPU decoder on oracle(libFLAC-procedure)-encoded streams of every kind."""
    for name, (x, bps) in gen_cases().items():
        for level in (0, 5, 8):
            enc, fs = fo.encode(x, bps, 48000, level)
            hdr = flacfmt.parse_header(enc)
            try:
                out = nat.host_decode(enc[hdr.first_frame_offset:], x.shape[1], bps, 4096, 48000, x.shape[0])
                report(f"decode oracle-encoded {name} L{level}", np.array_equal(out, x))
            except Exception as e:  # noqa: BLE001
                report(f"decode oracle-encoded {name} L{level}", False, repr(e))


def check_normalize():
    import torch
    from oracle import normalization_oracle as no
    z = np.load(G / "normalization_vectors.npz")
    keys = sorted(set(k.rsplit("__", 1)[0] for k in z.files if k.endswith("__in")))
    L = nat.lib()
    for k in keys:
        x, a, b, p = z[k + "__in"], z[k + "__audio"], z[k + "__back"], z[k + "__params"]
        dt = k.split("__")[0]
        bits = int(p[2])
        xt = torch.from_numpy(x.view(np.uint8)).cuda()
        mm = torch.empty(2, dtype=torch.float64, device="cuda")
        nat.check(L.frb_minmax_flat(xt.data_ptr(), nat.DTYPE_CODES[dt], x.size, mm.data_ptr(), None), "minmax")
        mmh = mm.cpu().numpy()
        ok_mm = (mmh[0] == p[0] or (np.isnan(mmh[0]) and np.isnan(p[0]))) and (mmh[1] == p[1] or (np.isnan(mmh[1]) and np.isnan(p[1])))
        out16 = a.dtype == np.int16
        at = torch.empty(x.size, dtype=torch.int16 if out16 else torch.int32, device="cuda")
        nat.check(L.frb_normalize_flat(xt.data_ptr(), nat.DTYPE_CODES[dt], x.size, float(p[0]), float(p[1]), bits, at.data_ptr(), int(out16), None), "normalize")
        ok_n = np.array_equal(at.cpu().numpy().reshape(a.shape), a)
        ot = torch.empty(x.size * x.itemsize, dtype=torch.uint8, device="cuda")
        nat.check(L.frb_denormalize_flat(at.data_ptr(), 0 if out16 else 1, x.size, float(p[0]), float(p[1]), float(p[3]) if not out16 else 32767.0,
                                         ot.data_ptr(), nat.DTYPE_CODES[dt], None), "denormalize")
        torch.cuda.synchronize()
        back = ot.cpu().numpy().view(x.dtype).reshape(b.shape)
        ok_d = np.array_equal(back, b, equal_nan=True)
        report(f"normalize {k}", ok_mm and ok_n and ok_d, f"minmax={ok_mm} norm={ok_n} denorm={ok_d}")


if __name__ == "__main__":
    print("devices:", nat.require_cuda(), flush=True)
    steps = [check_decode_golden, check_oracle_streams_decode, check_encode, check_normalize]
    which = sys.argv[1:] or None
    for fn in steps:
        if which and fn.__name__ not in which:
            continue
        try:
            fn()
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            report(fn.__name__, False, "exception")
    bad = [n for n, ok in results if not ok]
    print(f"\n{len(results) - len(bad)}/{len(results)} passed; launches={nat.lib().frb_launch_count()}")
    for n in bad:
        print("  FAILED:", n)
    sys.exit(1 if bad else 0)
