// frb_tilemap.cuh -- per-tile sample mapping kernels (min/max, normalise, denormalise) for a planar
// (bands,H,W) raster resident in HBM.  Reference: normalization.py:149-187 and :222-249, applied per tile
// as cli.py:553-594 does.
//
// v1 (profiles/r01_launches_c3_v1.csv) spent 3.8 / 2.6 / 4.2 ms on C3 where the HBM floor is ~0.9 / 0.3 /
// 0.9 ms: a runtime dtype switch and an integer division per element, plus ~40 fp64 instructions per
// sample (two IEEE divisions) on a 64-lane fp64 pipe.  v2: kernels are templated on the element type,
// walk (band,row) pairs so no per-element division is needed, and replace the per-sample fp64 division
// by exact lookup tables whose entries are computed with the very same operations:
//   * normalise, 8/16-bit sources: per-tile table over the tile's value range [min,max] (when it spans
//     fewer than kNormLutCap values), entry = normalize_one(value);
//   * denormalise, 16-bit audio (scale 32767): one table t[a] = (a/32767 + 1)/2 shared by all tiles, then
//     x = t*range + min with the same two rounded operations.
#pragma once
#include "frb_normalize.cuh"

namespace frb {

constexpr uint32_t kNormLutCap = 16384;      // per-tile normalise table entries (int32)
constexpr uint32_t kDenormLutN = 65536;      // audio value + 32768

struct MapWorkspace {
    int32_t *norm_lut;       // n_tiles * kNormLutCap
    double *denorm_lut;      // kDenormLutN
};
static inline size_t map_ws_layout(uint32_t n_tiles, void *base, MapWorkspace *w) {
    size_t off = 0;
    uint8_t *b = (uint8_t *)base;
    if (w) w->denorm_lut = (double *)(b + off);
    off += (size_t)kDenormLutN * 8;
    if (w) w->norm_lut = (int32_t *)(b + off);
    off += (size_t)n_tiles * kNormLutCap * 4;
    return off + 256;
}

template <typename T> struct is_small_int { static constexpr bool value = false; };
template <> struct is_small_int<uint8_t> { static constexpr bool value = true; };
template <> struct is_small_int<int8_t> { static constexpr bool value = true; };
template <> struct is_small_int<uint16_t> { static constexpr bool value = true; };
template <> struct is_small_int<int16_t> { static constexpr bool value = true; };

// ---------------------------------------------------------------- min/max
template <typename T>
__global__ void __launch_bounds__(256)
k_minmax_tiles(const T *__restrict__ raster, uint32_t bands, uint32_t H, uint32_t W,
               const frb_tile *__restrict__ tiles, unsigned long long *keys) {
    const frb_tile t = tiles[blockIdx.y];
    const uint32_t rows = bands * t.h;
    bool any = false;
    double mn = 0.0, mx = 0.0;
    for (uint32_t ry = blockIdx.x; ry < rows; ry += gridDim.x) {
        const uint32_t c = ry / t.h, y = ry - c * t.h;
        const T *src = raster + ((size_t)c * H + (t.row_off + y)) * W + t.col_off;
        for (uint32_t x = threadIdx.x; x < t.w; x += blockDim.x) {
            const double v = (double)src[x];
            if (v == v) {
                if (!any) { mn = mx = v; any = true; }
                else { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
            }
        }
    }
    block_minmax_commit(any ? dkey(mn) : kKeyMinInit, any ? dkey(mx) : kKeyMaxInit, keys + 2 * (size_t)blockIdx.y);
}

// ---------------------------------------------------------------- normalise
template <typename T>
__global__ void __launch_bounds__(256)
k_build_norm_lut(const double *__restrict__ minmax, int bits, int32_t *__restrict__ lut) {
    const double mn = minmax[2 * blockIdx.y], mx = minmax[2 * blockIdx.y + 1];
    if (!(mx - mn < (double)kNormLutCap)) return;          // also false for NaN: those tiles use the direct path
    const double range = (mx <= mn) ? 1.0 : __dsub_rn(mx, mn);
    const double scale = scale_for_bits(bits);
    const uint32_t cnt = (uint32_t)(mx - mn) + 1;
    int32_t *dst = lut + (size_t)blockIdx.y * kNormLutCap;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x)
        dst[j] = normalize_one(__dadd_rn(mn, (double)j), mn, range, scale);     // mn + j is exact (small integers)
}

template <typename T>
__global__ void __launch_bounds__(256)
k_normalize_tiles(const T *__restrict__ raster, uint32_t bands, uint32_t H, uint32_t W,
                  const frb_tile *__restrict__ tiles, const double *__restrict__ minmax, int bits,
                  int32_t *__restrict__ audio, const int64_t *__restrict__ audio_base,
                  const int32_t *__restrict__ lut_all) {
    const frb_tile t = tiles[blockIdx.y];
    const uint32_t n = t.h * t.w;
    const double mn = minmax[2 * blockIdx.y], mx = minmax[2 * blockIdx.y + 1];
    const double range = (mx <= mn) ? 1.0 : __dsub_rn(mx, mn);
    const double scale = scale_for_bits(bits);
    int32_t *dst = audio + audio_base[blockIdx.y];
    const uint32_t rows = bands * t.h;
    const bool use_lut = is_small_int<T>::value && lut_all != nullptr && (mx - mn < (double)kNormLutCap);
    const int32_t *lut = lut_all ? lut_all + (size_t)blockIdx.y * kNormLutCap : nullptr;
    const int32_t vmin = use_lut ? (int32_t)mn : 0;
    for (uint32_t ry = blockIdx.x; ry < rows; ry += gridDim.x) {
        const uint32_t c = ry / t.h, y = ry - c * t.h;
        const T *src = raster + ((size_t)c * H + (t.row_off + y)) * W + t.col_off;
        int32_t *out = dst + (size_t)c * n + (size_t)y * t.w;
        if (use_lut) {
            for (uint32_t x = threadIdx.x; x < t.w; x += blockDim.x) out[x] = __ldg(lut + ((int32_t)src[x] - vmin));
        } else {
            for (uint32_t x = threadIdx.x; x < t.w; x += blockDim.x) out[x] = normalize_one((double)src[x], mn, range, scale);
        }
    }
}

// ---------------------------------------------------------------- denormalise
__global__ void __launch_bounds__(256)
k_build_denorm_lut(double scale, double *__restrict__ lut) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < kDenormLutN) {
        const double a = (double)((int32_t)j - 32768);
        lut[j] = __ddiv_rn(__dadd_rn(__ddiv_rn(a, scale), 1.0), 2.0);      // (a/scale + 1)/2, same ops as denormalize_one
    }
}

template <typename T> __device__ __forceinline__ T denorm_cast(double v) { return cast_round_out<T>(v); }
template <> __device__ __forceinline__ float denorm_cast<float>(double v) { return __double2float_rn(v); }
template <> __device__ __forceinline__ double denorm_cast<double>(double v) { return v; }

template <typename T>
__global__ void __launch_bounds__(256)
k_denormalize_tiles(const int32_t *__restrict__ audio, const int64_t *__restrict__ audio_base,
                    const frb_tile *__restrict__ tiles, const double *__restrict__ minmax, double scale,
                    T *__restrict__ raster, uint32_t bands, uint32_t H, uint32_t W, const double *__restrict__ lut) {
    const frb_tile t = tiles[blockIdx.y];
    const uint32_t n = t.h * t.w;
    const double mn = minmax[2 * blockIdx.y], mx = minmax[2 * blockIdx.y + 1];
    const double range = __dsub_rn(mx, mn);                 // denormalize uses max-min unconditionally (:239)
    const int32_t *src = audio + audio_base[blockIdx.y];
    const uint32_t rows = bands * t.h;
    for (uint32_t ry = blockIdx.x; ry < rows; ry += gridDim.x) {
        const uint32_t c = ry / t.h, y = ry - c * t.h;
        T *out = raster + ((size_t)c * H + (t.row_off + y)) * W + t.col_off;
        const int32_t *in = src + (size_t)c * n + (size_t)y * t.w;
        for (uint32_t x = threadIdx.x; x < t.w; x += blockDim.x) {
            const int32_t a = in[x];
            double v;
            if (lut != nullptr && a >= -32768 && a <= 32767)
                v = __dadd_rn(__dmul_rn(__ldg(lut + (a + 32768)), range), mn);
            else
                v = denormalize_one((double)a, scale, mn, range);
            out[x] = denorm_cast<T>(v);
        }
    }
}

static inline dim3 tile_grid_dims(uint32_t n_tiles) {
    uint32_t per_tile = (kNumSMs * 8 + n_tiles - 1) / n_tiles;
    if (per_tile < 1) per_tile = 1;
    return dim3(per_tile, n_tiles);
}

#define FRB_DISPATCH_DTYPE(dtype, CALL)                                   \
    switch (dtype) {                                                      \
        case FRB_U8:  { using T = uint8_t;  CALL; } break;                \
        case FRB_I8:  { using T = int8_t;   CALL; } break;                \
        case FRB_U16: { using T = uint16_t; CALL; } break;                \
        case FRB_I16: { using T = int16_t;  CALL; } break;                \
        case FRB_U32: { using T = uint32_t; CALL; } break;                \
        case FRB_I32: { using T = int32_t;  CALL; } break;                \
        case FRB_F32: { using T = float;    CALL; } break;                \
        default:      { using T = double;   CALL; } break;                \
    }

}  // namespace frb

extern "C" int frb_sample_map_workspace_size(uint32_t n_tiles, size_t *bytes) {
    if (!bytes || !n_tiles) return FRB_ERR_INVALID_ARG;
    *bytes = frb::map_ws_layout(n_tiles, nullptr, nullptr);
    return FRB_OK;
}

extern "C" int frb_minmax_tiles(const void *d_raster, int dtype, uint32_t bands, uint32_t H, uint32_t W,
                                const frb_tile *d_tiles, uint32_t n_tiles, double *d_minmax, void *stream) {
    using namespace frb;
    if (!d_raster || !d_tiles || !d_minmax || dtype < 0 || dtype > FRB_F64 || !bands || !n_tiles) return FRB_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    k_minmax_init<<<(n_tiles + 255) / 256, 256, 0, s>>>((unsigned long long *)d_minmax, n_tiles);
    FRB_LAUNCH_CHECK("k_minmax_init");
    const dim3 grid = tile_grid_dims(n_tiles);
    FRB_DISPATCH_DTYPE(dtype, (k_minmax_tiles<T><<<grid, 256, 0, s>>>((const T *)d_raster, bands, H, W, d_tiles, (unsigned long long *)d_minmax)));
    FRB_LAUNCH_CHECK("k_minmax_tiles");
    k_minmax_finish<<<(n_tiles + 255) / 256, 256, 0, s>>>((unsigned long long *)d_minmax, n_tiles);
    FRB_LAUNCH_CHECK("k_minmax_finish");
    return FRB_OK;
}

extern "C" int frb_normalize_tiles(const void *d_raster, int dtype, uint32_t bands, uint32_t H, uint32_t W,
                                   const frb_tile *d_tiles, uint32_t n_tiles, const double *d_minmax,
                                   int bits_per_sample, int32_t *d_audio, const int64_t *d_audio_base,
                                   void *d_workspace, size_t workspace_bytes, void *stream) {
    using namespace frb;
    if (!d_raster || !d_tiles || !d_minmax || !d_audio || !d_audio_base || dtype < 0 || dtype > FRB_F64 || !bands || !n_tiles)
        return FRB_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int32_t *lut = nullptr;
    if (d_workspace && dtype <= FRB_I16) {
        MapWorkspace w;
        if (map_ws_layout(n_tiles, d_workspace, &w) > workspace_bytes) return FRB_ERR_OVERFLOW;
        k_build_norm_lut<int><<<dim3(8, n_tiles), 256, 0, s>>>(d_minmax, bits_per_sample, w.norm_lut);
        FRB_LAUNCH_CHECK("k_build_norm_lut");
        lut = w.norm_lut;
    }
    const dim3 grid = tile_grid_dims(n_tiles);
    FRB_DISPATCH_DTYPE(dtype, (k_normalize_tiles<T><<<grid, 256, 0, s>>>((const T *)d_raster, bands, H, W, d_tiles, d_minmax,
                                                                        bits_per_sample, d_audio, d_audio_base, lut)));
    FRB_LAUNCH_CHECK("k_normalize_tiles");
    return FRB_OK;
}

extern "C" int frb_denormalize_tiles(const int32_t *d_audio, const int64_t *d_audio_base,
                                     const frb_tile *d_tiles, uint32_t n_tiles, const double *d_minmax,
                                     double scale, void *d_raster, int dtype, uint32_t bands, uint32_t H,
                                     uint32_t W, void *d_workspace, size_t workspace_bytes, void *stream) {
    using namespace frb;
    if (!d_raster || !d_tiles || !d_minmax || !d_audio || !d_audio_base || dtype < 0 || dtype > FRB_F64 || !bands || !n_tiles)
        return FRB_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const double *lut = nullptr;
    if (d_workspace && scale == 32767.0) {
        MapWorkspace w;
        if (map_ws_layout(n_tiles, d_workspace, &w) > workspace_bytes) return FRB_ERR_OVERFLOW;
        k_build_denorm_lut<<<kDenormLutN / 256, 256, 0, s>>>(scale, w.denorm_lut);
        FRB_LAUNCH_CHECK("k_build_denorm_lut");
        lut = w.denorm_lut;
    }
    const dim3 grid = tile_grid_dims(n_tiles);
    FRB_DISPATCH_DTYPE(dtype, (k_denormalize_tiles<T><<<grid, 256, 0, s>>>(d_audio, d_audio_base, d_tiles, d_minmax, scale,
                                                                          (T *)d_raster, bands, H, W, lut)));
    FRB_LAUNCH_CHECK("k_denormalize_tiles");
    return FRB_OK;
}
