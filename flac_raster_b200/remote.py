"""Remote byte-range access (reference src/flac_raster/remote.py).

The names the package surface exports are kept: is_remote_url, RemoteFile.read_range,
open_remote, read_remote_range, download_remote.  HTTP(S) uses Range requests
(remote.py:153-168); cloud schemes need obstore, as in the reference.

Added for the tile path: read_range_into writes a range straight into a caller
buffer -- SpatialFLACStreamer hands it slices of its pinned staging buffer, several ranges in
flight on a thread pool (one keep-alive connection per thread), and sends every piece to the
GPU as soon as it has arrived, so the fetches overlap the host-to-device copies.
"""
from __future__ import annotations

import http.client
import tempfile
import threading
import urllib.parse
import urllib.request
from pathlib import Path
from typing import Optional, Union

_SCHEMES = ("http://", "https://", "s3://", "gs://", "az://", "abfs://", "abfss://")


def is_remote_url(path: Union[str, Path]) -> bool:
    return str(path).lower().startswith(_SCHEMES)


def get_url_scheme(url: str) -> str:
    return url.split("://", 1)[0].lower() if "://" in url else ""


class RemoteFile:
    """Byte-range reader with the reference's interface (remote.py:61-204)."""

    def __init__(self, url: str):
        self.url = url
        self.scheme = get_url_scheme(url)
        self._store = None
        self._path = None
        if self.scheme not in ("http", "https"):
            try:
                import obstore  # type: ignore  # noqa: F401
                from obstore.store import from_url  # type: ignore
            except ImportError as e:
                raise ImportError("obstore is required for s3://, gs:// and az:// URLs") from e
            base, _, key = url.rpartition("/")
            self._store, self._path = from_url(base), key

    def read_range(self, start: int, end: int) -> bytes:
        """Inclusive byte range, as the reference (remote.py:137-177)."""
        if self._store is not None:
            import obstore  # type: ignore

            return bytes(obstore.get_range(self._store, self._path, start=start, end=end + 1))
        req = urllib.request.Request(self.url, headers={"Range": f"bytes={start}-{end}"})
        with urllib.request.urlopen(req, timeout=30) as r:
            data = r.read()
        if len(data) > end - start + 1:      # server ignored the Range header
            data = data[start:end + 1]
        return data

    def read_range_into(self, start: int, end: int, out) -> int:
        """Inclusive byte range written into the writable buffer `out` (len >= end - start + 1); returns the byte count.
        HTTP(S): one keep-alive connection per calling thread, the body is received straight into `out`."""
        want = end - start + 1
        mv = memoryview(out).cast("B")
        if self._store is not None or self.scheme not in ("http", "https"):
            data = self.read_range(start, end)
            mv[:len(data)] = data
            return len(data)
        u = urllib.parse.urlsplit(self.url)
        path = (u.path or "/") + ("?" + u.query if u.query else "")
        local = _conn_cache.__dict__.setdefault("conns", {})
        key = (u.scheme, u.netloc)
        for attempt in (0, 1):
            conn = local.get(key)
            if conn is None:
                cls = http.client.HTTPSConnection if u.scheme == "https" else http.client.HTTPConnection
                conn = local[key] = cls(u.netloc, timeout=30)
            try:
                conn.request("GET", path, headers={"Range": f"bytes={start}-{end}", "Connection": "keep-alive"})
                r = conn.getresponse()
                if r.status == 206:
                    got = 0
                    while got < want:
                        n = r.readinto(mv[got:want])
                        if not n:
                            break
                        got += n
                    r.read()                       # drain (nothing left on a conforming server) so the connection can be reused
                    return got
                body = r.read()                    # 200: the server ignored the Range header (remote.py:166-168)
                if r.status != 200:
                    raise OSError(f"HTTP {r.status} for {self.url}")
                data = body[start:end + 1]
                mv[:len(data)] = data
                return len(data)
            except (http.client.HTTPException, ConnectionError, OSError):
                local.pop(key, None)
                try:
                    conn.close()
                except Exception:  # noqa: BLE001
                    pass
                if attempt:
                    raise
        return 0

    def read_all(self) -> bytes:
        if self._store is not None:
            import obstore  # type: ignore

            return bytes(obstore.get(self._store, self._path).bytes())
        with urllib.request.urlopen(self.url, timeout=60) as r:
            return r.read()

    def download_to_temp(self) -> Path:
        suffix = Path(self.url.split("?")[0]).suffix or ".bin"
        with tempfile.NamedTemporaryFile(suffix=suffix, delete=False) as tmp:
            tmp.write(self.read_all())
            return Path(tmp.name)


_conn_cache = threading.local()


def open_remote(url: str) -> RemoteFile:
    return RemoteFile(url)


def read_remote_range(url: str, start: int, end: int) -> bytes:
    return RemoteFile(url).read_range(start, end)


def download_remote(url: str, output_path: Optional[Path] = None) -> Path:
    rf = RemoteFile(url)
    if output_path is None:
        return rf.download_to_temp()
    Path(output_path).write_bytes(rf.read_all())
    return Path(output_path)
