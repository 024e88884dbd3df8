// frb_normalize.cuh -- sample mapping kernels (normalization.py:126-253 of the reference).
//
// All arithmetic is fp64 with explicit round-to-nearest intrinsics in numpy's
// operation order so results are bit-identical to the reference:
//   n = 2.0*(x - min)/range - 1.0 ; clip[-1,1] ; NaN->0 ; trunc(n*scale)
// HBM-bound elementwise kernels: grid-stride, coalesced, 148*k CTAs.
#pragma once
#include "frb_common.cuh"

namespace frb {

__device__ __forceinline__ double load_as_double(const void *p, int dtype, uint64_t i) {
    switch (dtype) {
        case FRB_U8:  return (double)((const uint8_t *)p)[i];
        case FRB_I8:  return (double)((const int8_t *)p)[i];
        case FRB_U16: return (double)((const uint16_t *)p)[i];
        case FRB_I16: return (double)((const int16_t *)p)[i];
        case FRB_U32: return (double)((const uint32_t *)p)[i];
        case FRB_I32: return (double)((const int32_t *)p)[i];
        case FRB_F32: return (double)((const float *)p)[i];
        default:      return ((const double *)p)[i];
    }
}

// order-preserving map double -> uint64 (for atomicMin/Max); NaN never passed in
__device__ __forceinline__ unsigned long long dkey(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dunkey(unsigned long long k) {
    unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
constexpr unsigned long long kKeyMinInit = 0xFFFFFFFFFFFFFFFFull;   // > every key
constexpr unsigned long long kKeyMaxInit = 0ull;                    // < every key

__device__ __forceinline__ int32_t normalize_one(double x, double mn, double range, double scale) {
    double n = __dsub_rn(__ddiv_rn(__dmul_rn(2.0, __dsub_rn(x, mn)), range), 1.0);
    // np.clip == minimum(maximum(n,-1),1) (NaN propagates), then NaN -> 0
    if (n != n) n = 0.0;
    else { n = n < -1.0 ? -1.0 : n; n = n > 1.0 ? 1.0 : n; }
    return __double2int_rz(__dmul_rn(n, scale));
}

template <typename T>
__device__ __forceinline__ T cast_round_out(double v);
template <> __device__ __forceinline__ uint8_t  cast_round_out<uint8_t>(double v)  { return (uint8_t)(int32_t)__double2ll_rn(v); }
template <> __device__ __forceinline__ int8_t   cast_round_out<int8_t>(double v)   { return (int8_t)(int32_t)__double2ll_rn(v); }
template <> __device__ __forceinline__ uint16_t cast_round_out<uint16_t>(double v) { return (uint16_t)(int32_t)__double2ll_rn(v); }
template <> __device__ __forceinline__ int16_t  cast_round_out<int16_t>(double v)  { return (int16_t)(int32_t)__double2ll_rn(v); }
template <> __device__ __forceinline__ uint32_t cast_round_out<uint32_t>(double v) { return (uint32_t)__double2ll_rn(v); }
template <> __device__ __forceinline__ int32_t  cast_round_out<int32_t>(double v)  { return (int32_t)__double2ll_rn(v); }

__device__ __forceinline__ void store_denorm(void *p, int dtype, uint64_t i, double v) {
    // integer dtypes: np.round (half to even) then cast; float dtypes: plain cast
    switch (dtype) {
        case FRB_U8:  ((uint8_t *)p)[i]  = cast_round_out<uint8_t>(v); break;
        case FRB_I8:  ((int8_t *)p)[i]   = cast_round_out<int8_t>(v); break;
        case FRB_U16: ((uint16_t *)p)[i] = cast_round_out<uint16_t>(v); break;
        case FRB_I16: ((int16_t *)p)[i]  = cast_round_out<int16_t>(v); break;
        case FRB_U32: ((uint32_t *)p)[i] = cast_round_out<uint32_t>(v); break;
        case FRB_I32: ((int32_t *)p)[i]  = cast_round_out<int32_t>(v); break;
        case FRB_F32: ((float *)p)[i]    = __double2float_rn(v); break;
        default:      ((double *)p)[i]   = v; break;
    }
}

__device__ __forceinline__ double denormalize_one(double a, double scale, double mn, double range) {
    double n = __ddiv_rn(a, scale);
    return __dadd_rn(__dmul_rn(__ddiv_rn(__dadd_rn(n, 1.0), 2.0), range), mn);
}

// a / scale for an INTEGER a and one of the three constant scales, without the generic division sequence:
// q0 = a * (1/scale), one residual correction with two FMAs (Markstein).  For these operands the result equals
// the correctly rounded quotient for EVERY integer a of the audio range: frb_selftest_division compares it with
// __ddiv_rn exhaustively (tests/test_gpu_parity.py::test_constant_division_is_exact).  The generic division
// made k_denormalize_tiles fp64-bound (1.56 ms on C3 where the HBM floor is 0.9 ms).
__device__ __forceinline__ double div_by_scale(double a, double scale, double rcp) {
    const double q0 = __dmul_rn(a, rcp);
    const double r = __fma_rn(-q0, scale, a);
    return __fma_rn(r, rcp, q0);
}
__device__ __forceinline__ double denormalize_one_fast(double a, double scale, double rcp, double mn, double range) {
    const double n = div_by_scale(a, scale, rcp);
    return __dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(n, 1.0), 0.5), range), mn);      // x/2 == x*0.5 exactly
}
__global__ void k_selftest_division(double scale, long long lo, long long hi, unsigned long long *mismatches) {
    const double rcp = __drcp_rn(scale);
    unsigned long long bad = 0;
    for (long long a = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; a <= hi; a += (long long)gridDim.x * blockDim.x) {
        const double want = __ddiv_rn((double)a, scale), got = div_by_scale((double)a, scale, rcp);
        if (__double_as_longlong(want) != __double_as_longlong(got)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ---------------------------------------------------------------- min/max
__global__ void k_minmax_init(unsigned long long *keys, uint32_t n_tiles) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tiles) { keys[2 * i] = kKeyMinInit; keys[2 * i + 1] = kKeyMaxInit; }
}

__device__ __forceinline__ void block_minmax_commit(unsigned long long kmin, unsigned long long kmax,
                                                    unsigned long long *dst) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long a = __shfl_xor_sync(0xFFFFFFFFu, kmin, o);
        unsigned long long b = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
    }
    __shared__ unsigned long long smin[32], smax[32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
    __syncthreads();
    if (warp == 0) {
        int nw = (blockDim.x + 31) >> 5;
        kmin = lane < nw ? smin[lane] : kKeyMinInit;
        kmax = lane < nw ? smax[lane] : kKeyMaxInit;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long a = __shfl_xor_sync(0xFFFFFFFFu, kmin, o);
            unsigned long long b = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
            kmin = a < kmin ? a : kmin;
            kmax = b > kmax ? b : kmax;
        }
        if (lane == 0) {
            if (kmin != kKeyMinInit) atomicMin(&dst[0], kmin);
            if (kmax != kKeyMaxInit) atomicMax(&dst[1], kmax);
        }
    }
}

__global__ void __launch_bounds__(256)
k_minmax_flat(const void *__restrict__ src, int dtype, uint64_t n, unsigned long long *keys) {
    unsigned long long kmin = kKeyMinInit, kmax = kKeyMaxInit;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        double v = load_as_double(src, dtype, i);
        if (v == v) {
            unsigned long long k = dkey(v);
            kmin = k < kmin ? k : kmin;
            kmax = k > kmax ? k : kmax;
        }
    }
    block_minmax_commit(kmin, kmax, keys);
}

// keys -> doubles in place (all-NaN / empty -> NaN, like np.nanmin's warning path)
__global__ void k_minmax_finish(unsigned long long *keys, uint32_t n_tiles) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tiles) {
        unsigned long long a = keys[2 * i], b = keys[2 * i + 1];
        double mn = (a == kKeyMinInit) ? __longlong_as_double(0x7FF8000000000000ll) : dunkey(a);
        double mx = (b == kKeyMaxInit) ? __longlong_as_double(0x7FF8000000000000ll) : dunkey(b);
        ((double *)keys)[2 * i] = mn;
        ((double *)keys)[2 * i + 1] = mx;
    }
}

// ---------------------------------------------------------------- normalise
__device__ __forceinline__ double scale_for_bits(int bits) {
    return bits == 16 ? 32767.0 : bits == 24 ? 8388607.0 : 2147483647.0;
}

__global__ void __launch_bounds__(256)
k_normalize_flat(const void *__restrict__ src, int dtype, uint64_t n, double mn, double mx, int bits,
                 void *__restrict__ out, int out16) {
    const double range = (mx <= mn) ? 1.0 : __dsub_rn(mx, mn);
    const double scale = scale_for_bits(bits);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        int32_t v = normalize_one(load_as_double(src, dtype, i), mn, range, scale);
        if (out16) ((int16_t *)out)[i] = (int16_t)v; else ((int32_t *)out)[i] = v;
    }
}

__global__ void __launch_bounds__(256)
k_denormalize_flat(const void *__restrict__ audio, int kind, uint64_t n, double mn, double mx, double scale,
                   void *__restrict__ out, int dtype) {
    const double range = __dsub_rn(mx, mn);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        double a = kind == 0 ? (double)((const int16_t *)audio)[i]
                 : kind == 1 ? (double)((const int32_t *)audio)[i] : ((const double *)audio)[i];
        store_denorm(out, dtype, i, denormalize_one(a, scale, mn, range));
    }
}

static inline uint32_t grid_for(uint64_t n, uint32_t per_cta, uint32_t max_ctas) {
    uint64_t g = (n + per_cta - 1) / per_cta;
    if (g < 1) g = 1;
    if (g > max_ctas) g = max_ctas;
    return (uint32_t)g;
}

}  // namespace frb

// ------------------------------------------------------------------ C ABI
extern "C" int frb_minmax_flat(const void *d_src, int dtype, uint64_t n, double *d_minmax, void *stream) {
    using namespace frb;
    if (!d_src || !d_minmax || dtype < 0 || dtype > FRB_F64) return FRB_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    k_minmax_init<<<1, 32, 0, s>>>((unsigned long long *)d_minmax, 1);
    FRB_LAUNCH_CHECK("k_minmax_init");
    k_minmax_flat<<<grid_for(n, 256 * 16, kNumSMs * 8), 256, 0, s>>>(d_src, dtype, n, (unsigned long long *)d_minmax);
    FRB_LAUNCH_CHECK("k_minmax_flat");
    k_minmax_finish<<<1, 32, 0, s>>>((unsigned long long *)d_minmax, 1);
    FRB_LAUNCH_CHECK("k_minmax_finish");
    return FRB_OK;
}

extern "C" int frb_selftest_division(double scale, int64_t lo, int64_t hi, uint64_t *h_mismatches, void *stream) {
    using namespace frb;
    if (!h_mismatches || hi < lo || !(scale > 0.0)) return FRB_ERR_INVALID_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *d = nullptr;
    FRB_CUDA(cudaMalloc(&d, 8));
    cudaError_t e = cudaMemsetAsync(d, 0, 8, s);
    if (e == cudaSuccess) {
        k_selftest_division<<<kNumSMs * 8, 256, 0, s>>>(scale, (long long)lo, (long long)hi, d);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaMemcpyAsync(h_mismatches, d, 8, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "frb_selftest_division");
    return FRB_OK;
}

extern "C" int frb_normalize_flat(const void *d_src, int dtype, uint64_t n, double data_min, double data_max,
                                  int bits_per_sample, void *d_out, int out16, void *stream) {
    using namespace frb;
    if (!d_src || !d_out || dtype < 0 || dtype > FRB_F64) return FRB_ERR_INVALID_ARG;
    if (n == 0) return FRB_OK;
    k_normalize_flat<<<grid_for(n, 256 * 8, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
        d_src, dtype, n, data_min, data_max, bits_per_sample, d_out, out16);
    FRB_LAUNCH_CHECK("k_normalize_flat");
    return FRB_OK;
}

extern "C" int frb_denormalize_flat(const void *d_audio, int audio_kind, uint64_t n, double data_min,
                                    double data_max, double scale, void *d_out, int dtype, void *stream) {
    using namespace frb;
    if (!d_audio || !d_out || dtype < 0 || dtype > FRB_F64 || audio_kind < 0 || audio_kind > 2) return FRB_ERR_INVALID_ARG;
    if (n == 0) return FRB_OK;
    k_denormalize_flat<<<grid_for(n, 256 * 8, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
        d_audio, audio_kind, n, data_min, data_max, scale, d_out, dtype);
    FRB_LAUNCH_CHECK("k_denormalize_flat");
    return FRB_OK;
}
