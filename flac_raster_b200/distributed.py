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


def allgather_tile_sizes(local_sizes: np.ndarray, n_tiles: int, rank: int, world: int, device=None) -> np.ndarray:
    """All ranks obtain the byte size of every tile (int64[n_tiles]) -- the one collective of the path."""
    if world == 1:
        return np.asarray(local_sizes, dtype=np.int64).copy()
    per = max(shard_range(n_tiles, r, world)[1] - shard_range(n_tiles, r, world)[0] for r in range(world))
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    send = torch.zeros(per, dtype=torch.int64, device=device)
    send[:len(local_sizes)] = torch.from_numpy(np.asarray(local_sizes, dtype=np.int64)).to(device)
    recv = torch.empty(per * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(world, per)
    out = np.zeros(n_tiles, dtype=np.int64)
    for r in range(world):
        a, b = shard_range(n_tiles, r, world)
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
                            payload: np.ndarray, offsets: np.ndarray, sizes: np.ndarray):
    """Rank 0 writes [u32][JSON]; every rank pwrite()s its tiles at header + byte_offset."""
    index_json = json.dumps(index, separators=(",", ":")).encode("utf-8")
    header_size = 4 + len(index_json)
    flags = os.O_WRONLY | os.O_CREAT
    fd = os.open(path, flags, 0o644)
    try:
        if rank == 0:
            os.pwrite(fd, len(index_json).to_bytes(4, "big") + index_json, 0)
        mv = memoryview(payload)
        for j, (h, o, s) in enumerate(zip(headers, offsets, sizes)):
            pos = header_size + index["frames"][first_tile + j]["byte_offset"]
            os.pwrite(fd, h, pos)
            os.pwrite(fd, mv[int(o):int(o) + int(s)], pos + len(h))
    finally:
        os.close(fd)
    return header_size


def encode_streaming_sharded(raster_dev, row_origin: int, full_shape: Tuple[int, int, int], transform, crs, nodata,
                             dtype_name: str, tile_size: int, compression_level: int, output_path: Optional[str],
                             rank: int, world: int, engine=None):
    """Each rank encodes its contiguous block of tiles of a (bands,H,W) raster.

    raster_dev holds (at least) the rows this rank's tiles touch, starting at global row
    `row_origin`.  Returns (index, local EncodedTiles, (first_tile, last_tile)).
    """
    from .engine import default_engine, tile_grid
    from .spatial_encoder import build_streaming_container, _tile_bbox

    eng = engine or default_engine()
    bands, H, W = full_shape
    tiles_all = tile_grid(H, W, tile_size)
    a, b = shard_range(len(tiles_all), rank, world)
    local = tiles_all[a:b].copy()
    local["row_off"] -= row_origin
    # per-tile transforms/bboxes are computed against the global grid
    gtrans = transform
    shifted = None
    if transform is not None:
        from .tiffio import window_transform
        shifted = window_transform(transform, 0, row_origin)
    index_local, headers, enc = build_streaming_container(raster_dev, shifted, crs, nodata, dtype_name, tile_size,
                                                          compression_level, tiles=local, engine=eng)
    file_sizes = np.array([f["byte_size"] for f in index_local["frames"]], dtype=np.int64)
    sizes_all = allgather_tile_sizes(file_sizes, len(tiles_all), rank, world)
    bboxes = [_tile_bbox(gtrans, int(t["col_off"]), int(t["row_off"]), int(t["w"]), int(t["h"]))[0] for t in tiles_all]
    template = {k: v for k, v in index_local.items() if k != "frames"}
    template.update({"width": W, "height": H})
    if transform is not None:
        template["transform"] = list(transform[:6]) + [0.0, 0.0, 1.0]
    index = build_index(template, tiles_all, bboxes, sizes_all)
    if output_path:
        payload = enc.payload.cpu().numpy()
        write_sharded_container(output_path, index, rank, a, headers, payload, enc.offsets, enc.sizes)
    return index, enc, (a, b)
