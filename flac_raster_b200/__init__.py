"""flac_raster_b200: B200-native (sm_100a CUDA) drop-in for flac-raster's codec hot path.

Exports every public name of the reference package surface (src/flac_raster/__init__.py:43-68):

    import flac_raster_b200 as flac_raster

The codec (normalisation, FLAC encode/decode) runs in hand-written CUDA kernels through the C ABI
in include/flacraster_b200.h; there is no CPU fallback.
"""
from .compare import compare_tiffs, display_comparison_table
from .converter import RasterFLACConverter
from .normalization import (
    NormalizationParams,
    calculate_audio_params,
    denormalize_from_audio,
    estimate_precision_loss,
    normalize_to_audio,
)
from .remote import download_remote, is_remote_url, open_remote
from .spatial_encoder import SpatialFLACEncoder, SpatialFLACStreamer, SpatialIndex

# optional async COG reader of the reference is outside the accelerated path
ASYNC_GEOTIFF_AVAILABLE = False
AsyncGeoTIFFReader = None
read_geotiff_async = None
read_tile_async = None

__version__ = "0.2.0"
__all__ = [
    "RasterFLACConverter",
    "compare_tiffs",
    "display_comparison_table",
    "SpatialFLACEncoder",
    "SpatialFLACStreamer",
    "SpatialIndex",
    "normalize_to_audio",
    "denormalize_from_audio",
    "calculate_audio_params",
    "NormalizationParams",
    "estimate_precision_loss",
    "is_remote_url",
    "open_remote",
    "download_remote",
    "ASYNC_GEOTIFF_AVAILABLE",
    "AsyncGeoTIFFReader",
    "read_geotiff_async",
    "read_tile_async",
]
