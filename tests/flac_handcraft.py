"""A tiny FLAC frame writer for decoder tests: builds streams that no libFLAC preset produces (LPC orders up to 32,
Rice2 parameters, escape-coded partitions, arbitrary partition orders) so that the GPU decoder can be checked against the
oracle decoder on every branch of RFC 9639's subframe syntax.  Test infrastructure only."""
from __future__ import annotations

import numpy as np


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value: int, n: int):
        for i in range(n - 1, -1, -1):
            self.bits.append((value >> i) & 1)

    def put_signed(self, value: int, n: int):
        self.put(value & ((1 << n) - 1), n)

    def unary(self, q: int):
        self.bits.extend([0] * q)
        self.bits.append(1)

    def pad(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def tobytes(self) -> bytes:
        assert len(self.bits) % 8 == 0
        return np.packbits(np.array(self.bits, dtype=np.uint8)).tobytes()


def utf8_number(v: int) -> bytes:
    if v < 0x80:
        return bytes([v])
    out = []
    n = 1
    while v >= (1 << (6 * n + (6 - n))):
        n += 1
    for _ in range(n):
        out.append(0x80 | (v & 0x3F))
        v >>= 6
    lead = (0xFF << (7 - n)) & 0xFF
    return bytes([lead | v] + out[::-1])


def lpc_subframe(w: BitWriter, x: np.ndarray, bps: int, order: int, coefs, shift: int, precision: int,
                 partition_order: int, rice2: bool, escape_partitions=()):
    """One LPC subframe of samples x (python ints).  Residuals are computed exactly; every partition takes the Rice
    parameter that suits its residuals, partitions listed in escape_partitions are stored as raw words."""
    n = len(x)
    res = []
    for i in range(order, n):
        pred = sum(int(coefs[j]) * int(x[i - 1 - j]) for j in range(order)) >> shift
        res.append(int(x[i]) - pred)
    w.put(0, 1); w.put(32 | (order - 1), 6); w.put(0, 1)             # header: LPC, no wasted bits
    for i in range(order):
        w.put_signed(int(x[i]), bps)
    w.put(precision - 1, 4); w.put_signed(shift, 5)
    for c in coefs:
        w.put_signed(int(c), precision)
    w.put(1 if rice2 else 0, 2); w.put(partition_order, 4)
    plen, esc = (5, 31) if rice2 else (4, 15)
    psize = n >> partition_order
    pos = 0
    for p in range(1 << partition_order):
        cnt = psize - (order if p == 0 else 0)
        part = res[pos:pos + cnt]; pos += cnt
        if p in escape_partitions:
            raw = max([max(v.bit_length() + 1 for v in part)] if part else [0])
            w.put(esc, plen); w.put(raw, 5)
            for v in part:
                w.put_signed(v, raw)
            continue
        fold = [(v << 1) if v >= 0 else ((-v) << 1) - 1 for v in part]
        mean = (sum(fold) // max(len(fold), 1)) if fold else 0
        k = min(max(mean.bit_length() - 1, 0), esc - 1)
        w.put(k, plen)
        for u in fold:
            w.unary(u >> k)
            if k:
                w.put(u & ((1 << k) - 1), k)


def frame(x: np.ndarray, bps: int, sample_rate_code: int, frame_number: int, blocksize_full: int, crc8, crc16, subs=None, **sub) -> bytes:
    """One frame of independent channels.  x.shape = (n,) with the subframe arguments as keywords, or (n, channels) with
    `subs` = one dict of lpc_subframe arguments per channel; sample_rate_code / bps are the header codes' meaning."""
    x = np.asarray(x)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
        subs = [sub]
    n, channels = x.shape
    bps_code = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}[bps]
    hdr = bytearray([0xFF, 0xF8])
    if n == blocksize_full and n in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
        bsc, extra = 8 + (n.bit_length() - 9), b""
    elif n <= 256:
        bsc, extra = 6, bytes([n - 1])
    else:
        bsc, extra = 7, bytes([(n - 1) >> 8, (n - 1) & 0xFF])
    hdr.append((bsc << 4) | sample_rate_code)
    hdr.append(((channels - 1) << 4) | (bps_code << 1))
    hdr += utf8_number(frame_number) + extra
    hdr.append(crc8(bytes(hdr)))
    w = BitWriter()
    for c in range(channels):
        lpc_subframe(w, x[:, c], bps, **subs[c])
    w.pad()
    body = bytes(hdr) + w.tobytes()
    c = crc16(body)
    return body + bytes([c >> 8, c & 0xFF])
