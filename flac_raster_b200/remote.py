"""Remote byte-range access (out of the accelerated path; reference src/flac_raster/remote.py).

Only the names the package surface exports are kept: is_remote_url, RemoteFile.read_range,
open_remote, read_remote_range, download_remote.  HTTP(S) uses urllib Range requests
(remote.py:153-168); cloud schemes need obstore, as in the reference.
"""
from __future__ import annotations

import tempfile
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
