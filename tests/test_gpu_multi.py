"""GPU suite: the multi-rank streaming-tile path through the PUBLIC API.

Two ranks of one torch.distributed group call SpatialFLACEncoder.encode on the same GeoTIFF: each codes its block of
tiles, the sizes are all-gathered, every rank pwrite()s its part of ONE container, which must be byte-identical to the
single-GPU file (cli.py:553-630).  The same for SpatialFLACStreamer.get_tiles_by_bbox, which splits the requested tiles
over the ranks.  With two GPUs the group is NCCL, one rank per GPU; on a one-GPU box both ranks share cuda:0 and the
size all-gather runs over gloo, so the path is covered there too."""
import os
import socket

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, backend, tif, out_path, res_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank if backend == "nccl" else 0)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from flac_raster_b200 import SpatialFLACEncoder, SpatialFLACStreamer
        idx = SpatialFLACEncoder(tile_size=100).encode(tif, out_path, streaming=True, compression_level=5)
        assert len(idx.frames) == 36                                      # 6x6 grid over 512x512 with 12-wide edge tiles
        s = SpatialFLACStreamer(out_path)
        t = s.metadata["transform"]
        res = s.get_tiles_by_bbox(t[2] - 1, t[5] + 600 * t[4], t[2] + 600 * t[0], t[5] + 1)     # everything: this rank's share
        np.savez(os.path.join(res_dir, f"rank{rank}.npz"), ids=np.array([m["frame_id"] for _, m in res]),
                 **{f"t{m['frame_id']}": a for a, m in res})
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_encode_and_bbox_decode_equal_single_gpu(tmp_path, world):
    import torch
    import torch.multiprocessing as mp
    from flac_raster_b200 import SpatialFLACEncoder
    from flac_raster_b200.tiffio import read_geotiff
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    tif = str(GOLDEN / "sample_dem.tif")
    single = tmp_path / "single.flac"
    SpatialFLACEncoder(tile_size=100).encode(tif, single, streaming=True, compression_level=5)
    sharded = tmp_path / "sharded.flac"
    sharded.write_bytes(b"\xEE" * (single.stat().st_size + 4096))        # stale longer file: must be truncated
    mp.spawn(_worker, args=(world, _free_port(), backend, tif, str(sharded), str(tmp_path)), nprocs=world, join=True)
    assert sharded.read_bytes() == single.read_bytes()
    src = read_geotiff(tif).data
    seen = []
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        ids = z["ids"].tolist()
        seen += ids
        for i in ids:
            ty, tx = divmod(i, 6)
            assert np.array_equal(z[f"t{i}"], src[:, ty * 100:(ty + 1) * 100, tx * 100:(tx + 1) * 100])
    assert seen == list(range(36))                                         # contiguous blocks in rank order, nothing twice
