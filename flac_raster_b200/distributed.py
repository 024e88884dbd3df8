"""Multi-GPU sharding of the streaming-tile path: one process per GPU, tiles split across ranks.

Tiles are independent complete FLAC files with per-tile min/max (cli.py:553-622), so the data
path needs no collective.  The only exchange is an all-gather of per-tile file sizes (8 bytes per
tile) whose exclusive scan gives every tile's byte_offset in the container index
(cli.py:615-621); each rank then pwrite()s its own tiles at those offsets and rank 0 writes
[u32 BE][JSON index].  NCCL over NVLink when the tensors live on the GPU, gloo on CPU (tests).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row-major block of tiles for `rank` (keeps each rank's bytes contiguous in the file)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_ranges_weighted(weights, world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks of items for ranks 0..world-1 whose largest weight sum is as small as a contiguous split allows.

    Tiles of a scene are not equally heavy: the last tile column and row are cut by the raster's edge (C3: 740 instead of
    1024 pixels wide or high), so equal tile COUNTS leave the rank with 16 full tiles 9 % above the mean at N = 8 while a
    split by pixel count comes within 2.4 %.  The step time of the sharded path is the slowest rank's.  Integer arithmetic
    only: every rank computes the same split from the tile grid.  Ranks past the last item get empty ranges."""
    w = [int(x) for x in weights]
    n = len(w)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world, 1) - 1)
    if any(x < 0 for x in w):
        raise ValueError("negative weight")

    def cut(cap: int) -> Optional[List[int]]:
        """greedy: fill every rank up to `cap`; block ends, or None when more than `world` blocks are needed"""
        ends, acc = [], 0
        for i, x in enumerate(w):
            if acc and acc + x > cap:
                ends.append(i)
                acc = 0
                if len(ends) >= world:
                    return None
            acc += x
        return ends + [n]

    lo, hi = max(max(w), -(-sum(w) // world)), sum(w)
    while lo < hi:                                         # smallest cap that needs no more than `world` blocks
        mid = (lo + hi) // 2
        if cut(mid) is None:
            lo = mid + 1
        else:
            hi = mid
    ends = cut(lo)
    # a heavy cap can leave ranks without work although there are items to give them: split the largest blocks further
    while len(ends) < min(world, n):
        spans = [(b - a, a, b) for a, b in zip([0] + ends[:-1], ends) if b - a > 1]
        _, a, b = max(spans, key=lambda t: (sum(w[t[1]:t[2]]), -t[1]))
        half, acc, m = sum(w[a:b]) / 2, 0, a
        while m < b - 1 and acc + w[m] <= half:
            acc += w[m]
            m += 1
        ends = sorted(ends + [max(m, a + 1)])
    starts = [0] + ends[:-1]
    out = list(zip(starts, ends))
    return out + [(n, n)] * (world - len(out))


def tile_shards(tiles_all: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """The split of a scene's tiles over `world` ranks: contiguous row-major blocks balanced by pixel count."""
    return shard_ranges_weighted(tiles_all["h"].astype(np.int64) * tiles_all["w"].astype(np.int64), world)


def init_from_env(backend: Optional[str] = None):
    """torchrun-style rendezvous (RANK/WORLD_SIZE/MASTER_ADDR/MASTER_PORT); returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process (and therefore its first-touch pinned host buffers) to the CPUs of the NUMA node its GPU hangs
    off.  With one process per GPU on a two-socket box the host side of the tile pipeline -- pinned staging, H2D of the
    raster, D2H of the frames -- otherwise crosses the socket interconnect for half the ranks.  Best effort: returns the
    node, or None when sysfs / NVML do not say."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip() != ""]
        idx = int(vis[device_index]) if vis and all(v.strip().isdigit() for v in vis) and device_index < len(vis) else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                      # NVML prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # noqa: BLE001
        return None


def allgather_tile_sizes(local_sizes: np.ndarray, n_tiles: int, rank: int, world: int, device=None, ranges=None) -> np.ndarray:
    """All ranks obtain the byte size of every tile (int64[n_tiles]) -- the one collective of the path.
    `ranges`: every rank's (first, last+1) tile (tile_shards); None = equal counts (shard_range)."""
    if world == 1:
        return np.asarray(local_sizes, dtype=np.int64).copy()
    if ranges is None:
        ranges = [shard_range(n_tiles, r, world) for r in range(world)]
    per = max(b - a for a, b in ranges)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    send = torch.zeros(per, dtype=torch.int64, device=device)
    send[:len(local_sizes)] = torch.from_numpy(np.asarray(local_sizes, dtype=np.int64)).to(device)
    recv = torch.empty(per * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(world, per)
    out = np.zeros(n_tiles, dtype=np.int64)
    for r, (a, b) in enumerate(ranges):
        out[a:b] = recv[r, :b - a]
    return out


class SizeExchange:
    """The path's one collective as a stream-ordered step of Engine.encode_tiles: all-gather of the per-tile frame sizes
    (int64, device to device over NCCL/NVLink) started behind the analysis kernels, running next to the frame assembly and
    joined in front of the step's single download, so that the other ranks' sizes arrive in the transfer the step needs anyway.  Ranks hold the contiguous
    blocks `ranges` (tile_shards), by default those of shard_range(n_tiles)."""

    def __init__(self, n_tiles: int, rank: int, world: int, device, ranges=None):
        self.n_tiles, self.rank, self.world, self.device = n_tiles, rank, world, device
        self.ranges = list(ranges) if ranges is not None else [shard_range(n_tiles, r, world) for r in range(world)]
        self.per = max(b - a for a, b in self.ranges)
        self.recv_count = self.per * world
        self._send = torch.zeros(self.per, dtype=torch.int64, device=device)

    def start(self, d_sizes: torch.Tensor, d_recv: torch.Tensor):
        """Enqueue the all-gather behind what the current stream holds (the analysis kernels that produce d_sizes) WITHOUT making
        the stream wait for it: the collective runs on the process group's own stream next to the frame assembly."""
        self._send[:d_sizes.numel()].copy_(d_sizes, non_blocking=True)
        self._work = dist.all_gather_into_tensor(d_recv, self._send, async_op=True)

    def wait(self):
        """Make the current stream wait for the collective started by start() (no host synchronisation)."""
        w, self._work = getattr(self, "_work", None), None
        if w is not None:
            w.wait()

    def enqueue(self, d_sizes: torch.Tensor, d_recv: torch.Tensor):
        self.start(d_sizes, d_recv)
        self.wait()

    def unpack(self, recv_host: np.ndarray) -> np.ndarray:
        recv = np.asarray(recv_host, dtype=np.int64).reshape(self.world, self.per)
        out = np.zeros(self.n_tiles, dtype=np.int64)
        for r, (a, b) in enumerate(self.ranges):
            out[a:b] = recv[r, :b - a]
        return out


def exclusive_scan(sizes: np.ndarray) -> np.ndarray:
    off = np.zeros(len(sizes), dtype=np.int64)
    np.cumsum(sizes[:-1], out=off[1:])
    return off


def build_index(template: Dict, tiles: np.ndarray, bboxes: List[List[float]], sizes: np.ndarray) -> Dict:
    """Container index (cli.py:538-547, :605-618) from the gathered sizes; identical on every rank."""
    offs = exclusive_scan(sizes)
    frames = []
    for i, t in enumerate(tiles):
        frames.append({
            "frame_id": i,
            "bbox": bboxes[i],
            "window": {"col_off": int(t["col_off"]), "row_off": int(t["row_off"]), "width": int(t["w"]), "height": int(t["h"])},
            "byte_offset": int(offs[i]),
            "byte_size": int(sizes[i]),
        })
    idx = dict(template)
    idx["frames"] = frames
    return idx


def write_sharded_container(path: str, index: Dict, rank: int, first_tile: int, headers: List[bytes],
                            payload, offsets: np.ndarray, sizes: np.ndarray, world: int = 1):
    """Rank 0 writes [u32][JSON]; every rank pwrite()s its tiles at header + byte_offset (cli.py:625-630).

    Rank 0 first creates the file and sets its length to header + sum(byte_size) -- re-encoding onto an existing,
    longer file must not leave stale bytes behind -- and only then (barrier) do the other ranks open and write it.
    A rank's tiles are contiguous in the file, so they go out as gathered writes (pwritev, <= 1024 buffers a call).
    `payload`: anything with the buffer protocol (numpy array, pinned tensor's .numpy())."""
    index_json = json.dumps(index, separators=(",", ":")).encode("utf-8")
    header_size = 4 + len(index_json)
    total = header_size + sum(int(f["byte_size"]) for f in index["frames"])
    multi = world > 1 and dist.is_available() and dist.is_initialized()
    if rank == 0:
        fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o644)
        try:
            os.ftruncate(fd, total)
            os.pwrite(fd, len(index_json).to_bytes(4, "big") + index_json, 0)
        except BaseException:
            os.close(fd)
            raise
    if multi:
        dist.barrier()
    if rank != 0:
        fd = os.open(path, os.O_WRONLY)
    try:
        mv = memoryview(payload)
        bufs: List = []
        pos0 = None
        expect = None
        for j, (h, o, s) in enumerate(zip(headers, offsets, sizes)):
            pos = header_size + int(index["frames"][first_tile + j]["byte_offset"])
            if pos0 is None:
                pos0 = expect = pos
            if pos != expect or len(bufs) >= 1022:       # not contiguous with the previous tile, or IOV_MAX reached
                _pwritev_all(fd, bufs, pos0)
                bufs, pos0 = [], pos
            bufs.append(h)
            bufs.append(mv[int(o):int(o) + int(s)])
            expect = pos + len(h) + int(s)
        if bufs:
            _pwritev_all(fd, bufs, pos0)
    finally:
        os.close(fd)
    if multi:
        dist.barrier()          # the container is complete on return, on every rank
    return header_size


def _pwritev_all(fd: int, bufs: List, pos: int):
    """os.pwritev until every byte is out (a gathered write may be short)."""
    bufs = [memoryview(b).cast("B") for b in bufs if len(b)]
    while bufs:
        n = os.pwritev(fd, bufs, pos)
        pos += n
        while bufs and n >= len(bufs[0]):
            n -= len(bufs[0])
            bufs.pop(0)
        if bufs and n:
            bufs[0] = bufs[0][n:]


def rows_of_shard(tiles_all: np.ndarray, a: int, b: int) -> Tuple[int, int]:
    """Raster rows [row0, row1) touched by tiles a..b-1 (a contiguous row-major block: at most a few tile rows)."""
    if b <= a:
        return 0, 0
    t = tiles_all[a:b]
    return int(t["row_off"].min()), int((t["row_off"].astype(np.int64) + t["h"]).max())


def shard_plan(height: int, width: int, tile_size: int, rank: int, world: int):
    """(all tiles, (first, last+1) tile of this rank, (row0, row1) raster rows this rank needs)."""
    from .engine import tile_grid

    tiles_all = tile_grid(height, width, tile_size)
    a, b = tile_shards(tiles_all, world)[rank]
    return tiles_all, (a, b), rows_of_shard(tiles_all, a, b)


def encode_streaming_sharded(raster, row_origin: int, full_shape: Tuple[int, int, int], transform, crs, nodata,
                             dtype_name: str, tile_size: int, compression_level: int, output_path: Optional[str],
                             rank: int, world: int, engine=None):
    """The streaming-tile loop of cli.py:553-630 with the tiles of ONE raster split over the ranks.

    Each rank encodes its contiguous row-major block of tiles (tile_shards: balanced by pixel count; one batched GPU call), the per-tile file sizes are
    all-gathered (the one collective; its exclusive scan is every tile's byte_offset, cli.py:615-621), rank 0
    writes [u32 BE][JSON index] and every rank pwrite()s its own tiles.  The file is byte-identical to the one
    SpatialFLACEncoder.encode writes on one GPU.

    raster: (bands, rows, W) device tensor, or host (ideally pinned) tensor, holding at least the rows this rank's
    tiles touch, starting at global row `row_origin` (see shard_plan).  Returns (index, local EncodedTiles,
    (first_tile, last_tile+1)).
    """
    from .engine import default_engine
    from .spatial_encoder import build_streaming_container, _tile_bbox
    eng = engine or default_engine()
    bands, H, W = full_shape
    tiles_all, (a, b), _ = shard_plan(H, W, tile_size, rank, world)
    if b > a:
        # tiles stay in global coordinates (their transform/bounds tags must equal the one-GPU file's bit for bit);
        # row_origin tells the engine where the local slab starts
        index_local, headers, enc = build_streaming_container(raster, transform, crs, nodata, dtype_name, tile_size,
                                                              compression_level, tiles=tiles_all[a:b], engine=eng,
                                                              row_origin=row_origin, full_shape=full_shape)
        file_sizes = np.array([f["byte_size"] for f in index_local["frames"]], dtype=np.int64)
    else:                                    # more ranks than tiles
        index_local, headers, enc, file_sizes = {}, [], None, np.zeros(0, dtype=np.int64)
    sizes_all = allgather_tile_sizes(file_sizes, len(tiles_all), rank, world, ranges=tile_shards(tiles_all, world))
    bboxes = [_tile_bbox(transform, int(t["col_off"]), int(t["row_off"]), int(t["w"]), int(t["h"]))[0] for t in tiles_all]
    template = {"crs": str(crs), "transform": (list(transform[:6]) + [0.0, 0.0, 1.0]) if transform else [],
                "width": W, "height": H, "bands": bands, "dtype": dtype_name, "tile_size": tile_size}
    index = build_index(template, tiles_all, bboxes, sizes_all)
    if output_path:
        if enc is None:
            payload, offs, szs = np.zeros(0, dtype=np.uint8), [], []
        elif enc.payload.is_cuda:
            # frames -> pinned staging (no pageable bounce buffer) -> pwritev
            n = int(enc.payload.numel())
            stage = eng._pinned("shard_out", n)
            stage[:n].copy_(enc.payload, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            payload, offs, szs = stage[:n].numpy(), enc.offsets, enc.sizes
        else:
            payload, offs, szs = enc.payload.numpy(), enc.offsets, enc.sizes
        write_sharded_container(output_path, index, rank, a, headers, payload, offs, szs, world)
    return index, enc, (a, b)


def current_rank_world() -> Tuple[int, int]:
    """(rank, world) of the initialised process group, (0, 1) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1
