import time, torch, numpy as np, sys
sys.path.insert(0, '.')
from flac_raster_b200.engine import Engine, tile_grid
from flac_raster_b200 import synth
dev = torch.device('cuda', 0)
eng = Engine(dev)
raster = synth.dem_int16_tiles(4096, 512, device=dev)
nb, H, W = raster.shape
tiles = tile_grid(H, W, 512)
host_in = torch.empty(raster.numel() * 2, dtype=torch.uint8).pin_memory()
host_in.copy_(raster.reshape(-1).view(torch.uint8))
host_raster = host_in.view(raster.dtype).reshape(raster.shape)
enc = eng.encode_tiles(raster, tiles, 5)
host_out = torch.empty(int(enc.sizes.sum()) + (1 << 20), dtype=torch.uint8).pin_memory()
import cProfile, pstats
for i in range(2):
    t0 = time.perf_counter(); e = eng.encode_tiles_host(host_raster, tiles, 5, host_out=host_out); torch.cuda.synchronize(); print('host path', time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable()
e = eng.encode_tiles_host(host_raster, tiles, 5, host_out=host_out); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
