import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle():
    from oracle import flac_oracle
    flac_oracle.build()
    return flac_oracle


@pytest.fixture(scope="session")
def rgb_pcm(oracle):
    pcm, _ = oracle.decode((GOLDEN / "sample_rgb.flac").read_bytes())
    return pcm


def signal_cases(seed=7):
    """Seeded (N,C) int32 signals with their FLAC bits per sample; shared by oracle and GPU tests."""
    rng = np.random.default_rng(seed)
    t = np.arange(50000)
    c = {}
    c["sine16_1ch"] = ((8000 * np.sin(t / 37.0) + 500 * rng.standard_normal(t.size)).astype(np.int32).reshape(-1, 1), 16)
    c["noise16_3ch"] = (rng.integers(-32768, 32767, size=(20000, 3)).astype(np.int32), 16)
    c["smooth16_8ch"] = (np.cumsum(rng.integers(-20, 21, size=(30000, 8)), axis=0).astype(np.int32), 16)
    c["const16"] = (np.full((9000, 1), -32767, dtype=np.int32), 16)
    c["zeros32"] = (np.zeros((8192, 1), dtype=np.int32), 32)
    c["wasted16_2ch"] = ((rng.integers(-2000, 2000, size=(12000, 2)) * 8).astype(np.int32), 16)
    c["dem24_1ch"] = ((4e6 * np.sin(t / 300.0) + 2000 * rng.standard_normal(t.size)).astype(np.int32).reshape(-1, 1), 32)
    c["noise24_2ch"] = (rng.integers(-8388607, 8388607, size=(10000, 2)).astype(np.int32), 32)
    c["tiny3"] = (np.array([[1], [2], [-3]], dtype=np.int32), 16)
    c["tail17"] = ((1000 * np.sin(np.arange(4096 + 17) / 9.0)).astype(np.int32).reshape(-1, 1), 16)
    c["exact8192"] = ((1000 * np.sin(np.arange(8192) / 9.0)).astype(np.int32).reshape(-1, 1), 16)
    c["extremes16"] = (np.where(rng.random((6000, 1)) < 0.5, -32767, 32767).astype(np.int32), 16)
    c["ramp24"] = ((np.arange(20000) * 397 - 4000000).astype(np.int32).reshape(-1, 1), 32)
    # two correlated channels: libFLAC's mid/side search (left/side, right/side and mid/side all win somewhere)
    left = np.cumsum(rng.integers(-60, 61, size=40000)).astype(np.int64)
    st = np.stack([left, left + rng.integers(-6, 7, size=left.size)], axis=1)
    st[12000:20000, 1] = rng.integers(-3000, 3000, size=8000)          # right turns into noise: independent / left-side
    st[26000:33000, 0] = -st[26000:33000, 1] + rng.integers(-4, 5, size=7000)   # anti-correlated: mid is small
    st[33000:, 0] = st[33000:, 1] + rng.integers(-40, 41, size=7000)                # left = smooth right + noise: right/side
    st[33000:, 1] = np.cumsum(rng.integers(-3, 4, size=7000)) + st[32999, 1]
    st[33000:, 0] = st[33000:, 1] + rng.integers(-40, 41, size=7000)
    c["stereo16_corr"] = (np.clip(st, -32767, 32767).astype(np.int32), 16)
    alt = np.where(np.arange(9000) % 2 == 0, 32000, -32000)
    c["stereo16_extreme"] = (np.stack([alt + rng.integers(-700, 700, size=9000), -alt + rng.integers(-700, 700, size=9000)],
                                      axis=1).astype(np.int32), 16)          # side = L-R swings over the full 17 bits
    return c


@pytest.fixture()
def range_http_server(tmp_path):
    """http://127.0.0.1:<port>/<name> serving files of tmp_path with Range support (206) -- or, for names starting with
    'norange_', ignoring the Range header (200 + whole body), the case the reference handles at remote.py:166-168.
    Yields (base_url, directory, log) where log collects the Range headers seen."""
    import http.server
    import threading

    log = []

    class H(http.server.BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def do_GET(self):
            p = tmp_path / self.path.lstrip("/").split("?")[0]
            if not p.is_file():
                self.send_response(404); self.send_header("Content-Length", "0"); self.end_headers()
                return
            data = p.read_bytes()
            rng = self.headers.get("Range")
            log.append(rng)
            if rng and not p.name.startswith("norange_"):
                a, b = rng.split("=")[1].split("-")
                a, b = int(a), min(int(b), len(data) - 1)
                body = data[a:b + 1]
                self.send_response(206)
                self.send_header("Content-Range", f"bytes {a}-{b}/{len(data)}")
            else:
                body = data
                self.send_response(200)
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, *a):
            pass

    srv = http.server.ThreadingHTTPServer(("127.0.0.1", 0), H)
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    try:
        yield f"http://127.0.0.1:{srv.server_address[1]}", tmp_path, log
    finally:
        srv.shutdown()
        srv.server_close()
