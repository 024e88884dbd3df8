"""CPU suite part 3: the N>1 path with world_size-2 gloo (size all-gather, offsets, sharded writes)."""
import json
import os
import socket
import struct

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, n_tiles):
    import torch.distributed as dist
    from flac_raster_b200 import _native as nat
    from flac_raster_b200.distributed import (allgather_tile_sizes, build_index, shard_ranges_weighted, write_sharded_container)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # unequal tiles: the pixel-count split gives the ranks different tile COUNTS (7 + 4 here), as the ragged edge of a scene does
    ranges = shard_ranges_weighted([1, 1, 1, 1, 1, 1, 3, 3, 3, 3, 1], world)
    assert ranges == [(0, 7), (7, 11)]
    a, b = ranges[rank]
    # fake per-tile files: header of 10 bytes + payload of (7 + tile id) bytes filled with the tile id
    headers = [bytes([0xA0 + (t % 16)]) * 10 for t in range(a, b)]
    psizes = np.array([7 + t for t in range(a, b)], dtype=np.int64)
    poffs = np.zeros(b - a, dtype=np.int64)
    np.cumsum(psizes[:-1], out=poffs[1:])
    payload = np.concatenate([np.full(7 + t, t, dtype=np.uint8) for t in range(a, b)])
    local_file_sizes = psizes + 10
    sizes = allgather_tile_sizes(local_file_sizes, n_tiles, rank, world, ranges=ranges)
    assert list(sizes) == [17 + t for t in range(n_tiles)]
    # the same exchange as the encode step runs it: started asynchronously behind the analysis, joined before the download
    import torch
    from flac_raster_b200.distributed import SizeExchange
    x = SizeExchange(n_tiles, rank, world, "cpu", ranges=ranges)
    recv = torch.zeros(x.recv_count, dtype=torch.int64)
    x.start(torch.from_numpy(local_file_sizes.copy()), recv)
    x.wait()
    assert list(x.unpack(recv.numpy())) == [17 + t for t in range(n_tiles)]
    x.enqueue(torch.from_numpy(local_file_sizes.copy()), recv)          # the blocking form
    assert list(x.unpack(recv.numpy())) == [17 + t for t in range(n_tiles)]
    tiles = np.zeros(n_tiles, dtype=nat.TILE_DTYPE)
    tiles["h"] = 1
    tiles["w"] = 1
    tiles["col_off"] = np.arange(n_tiles)
    index = build_index({"crs": "None", "transform": [], "width": n_tiles, "height": 1, "bands": 1, "dtype": "uint8", "tile_size": 1},
                        tiles, [[float(t), 0.0, float(t + 1), 1.0] for t in range(n_tiles)], sizes)
    write_sharded_container(path, index, rank, a, headers, payload, poffs, psizes, world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_container(tmp_path):
    n_tiles, world = 11, 2
    path = str(tmp_path / "sharded.flac")
    with open(path, "wb") as fh:                          # a stale, longer file from an earlier run must not leave bytes behind
        fh.write(b"\xEE" * 4096)
    mp.spawn(_worker, args=(world, _free_port(), path, n_tiles), nprocs=world, join=True)
    blob = open(path, "rb").read()
    (n,) = struct.unpack(">I", blob[:4])
    index = json.loads(blob[4:4 + n])
    hs = 4 + n
    assert [f["frame_id"] for f in index["frames"]] == list(range(n_tiles))
    off = 0
    for t, f in enumerate(index["frames"]):
        assert f["byte_offset"] == off and f["byte_size"] == 17 + t
        rec = blob[hs + off: hs + off + f["byte_size"]]
        assert rec[:10] == bytes([0xA0 + (t % 16)]) * 10 and rec[10:] == bytes([t]) * (7 + t)
        off += f["byte_size"]
    assert len(blob) == hs + off
