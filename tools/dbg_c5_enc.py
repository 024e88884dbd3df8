import time, torch, numpy as np, sys, cProfile, pstats
sys.path.insert(0, '.')
from flac_raster_b200.engine import Engine, tile_grid
from flac_raster_b200 import synth
dev = torch.device('cuda', 0)
eng = Engine(dev)
raster = synth.dem_int16_tiles(4096, 512, device=dev)
nb, H, W = raster.shape
tiles = tile_grid(H, W, 512)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); enc = eng.encode_tiles(raster, tiles, 5); torch.cuda.synchronize(); print('encode_tiles', (time.perf_counter() - t0) * 1e3, 'ms')
pr = cProfile.Profile(); pr.enable()
enc = eng.encode_tiles(raster, tiles, 5); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(16)
