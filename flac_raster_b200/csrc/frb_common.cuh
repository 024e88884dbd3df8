// frb_common.cuh -- shared helpers for the sm_100a FLAC engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string.h>
#include "../../include/flacraster_b200.h"

namespace frb {

// Unity build: this header is included exactly once by flacraster_b200.cu.
static thread_local char g_last_cuda_error[256] = "";
static std::atomic<uint64_t> g_launches{0};

inline int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_last_cuda_error, sizeof g_last_cuda_error, "%s: %s", what, cudaGetErrorString(e));
    return FRB_ERR_CUDA;
}

#define FRB_CUDA(call)                                                    \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return frb::cuda_fail(e__, #call);        \
    } while (0)

#define FRB_LAUNCH_CHECK(name)                                            \
    do {                                                                  \
        frb::g_launches.fetch_add(1, std::memory_order_relaxed);          \
        cudaError_t e__ = cudaGetLastError();                             \
        if (e__ != cudaSuccess) return frb::cuda_fail(e__, name);         \
    } while (0)

constexpr int kNumSMs = 148;   // B200

// cudaFuncSetAttribute is per device: returns true the first time it is called for the current device
inline bool first_call_on_device(bool (&done)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
}

// ---- small host <-> device transfers that stay off the copy engines -----------------------------------------------
// A bulk cudaMemcpyAsync in flight (the tile pipeline moves ~100-200 MB per stage and direction) occupies a DMA engine
// for milliseconds, and every small copy enqueued after it -- a 2 KB stream table, an 8-byte size read-back -- queues
// behind it: tools/e2e_timeline.py showed each stage's encode starting exactly when the previous stage's 130 MB
// device-to-host copy ended, i.e. copy and compute serialised (49 ms per C3 step against a 42 ms PCIe bound).  Small
// transfers therefore go through pinned (UVA-mapped) host memory and a copy KERNEL: SM loads/stores over PCIe, no DMA
// queue.  Memsets of small device ranges use a fill kernel for the same reason.
__global__ void k_copy_small(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, size_t bytes) {
    const size_t words = bytes >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    if (blockIdx.x == 0 && threadIdx.x < (bytes & 3)) ((uint8_t *)dst)[(words << 2) + threadIdx.x] = ((const uint8_t *)src)[(words << 2) + threadIdx.x];
}
__global__ void k_fill_small(uint32_t *__restrict__ dst, uint32_t v, size_t words) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
struct Mailbox {                       // per host thread: a ring of pinned memory; a slot lives until the ring wraps (32 MiB later)
    uint8_t *base = nullptr;
    size_t cap = 0, pos = 0;
    uint8_t *take(size_t bytes) {
        if (!base) {
            if (cudaMallocHost((void **)&base, kCap) != cudaSuccess) { base = nullptr; cudaGetLastError(); return nullptr; }
            cap = kCap;
        }
        bytes = (bytes + 63) & ~(size_t)63;
        if (bytes > cap / 4) return nullptr;               // not "small": the caller uses the DMA path
        if (pos + bytes > cap) {
            // wrapping onto slots whose copy kernels may, in principle, still be queued (a caller that never synchronises):
            // drain the device once per 32 MiB of small transfers
            cudaDeviceSynchronize();
            pos = 0;
        }
        uint8_t *p = base + pos;
        pos += bytes;
        return p;
    }
    static constexpr size_t kCap = 32u << 20;
};
static thread_local Mailbox t_mailbox;
static inline uint32_t small_grid(size_t bytes) { size_t g = (bytes / 4 + 255) / 256; return (uint32_t)(g < 1 ? 1 : g > 256 ? 256 : g); }

// host -> device, asynchronous on `s` (the source may be reused as soon as the call returns)
inline int small_upload(void *d_dst, const void *h_src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return FRB_OK;
    uint8_t *slot = ((reinterpret_cast<uintptr_t>(d_dst) & 3u) == 0) ? t_mailbox.take(bytes) : nullptr;
    if (!slot) {
        cudaError_t e = cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s);
        return e == cudaSuccess ? FRB_OK : cuda_fail(e, "cudaMemcpyAsync(H2D)");
    }
    memcpy(slot, h_src, bytes);
    k_copy_small<<<small_grid(bytes), 256, 0, s>>>((uint32_t *)d_dst, (const uint32_t *)slot, bytes);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FRB_OK : cuda_fail(e, "k_copy_small");
}
// device -> host; returns after the stream has been synchronised and h_dst holds the data
inline int small_download(void *h_dst, const void *d_src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return FRB_OK;
    uint8_t *slot = ((reinterpret_cast<uintptr_t>(d_src) & 3u) == 0) ? t_mailbox.take(bytes) : nullptr;
    cudaError_t e;
    if (!slot) {
        e = cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        return e == cudaSuccess ? FRB_OK : cuda_fail(e, "cudaMemcpyAsync(D2H)");
    }
    k_copy_small<<<small_grid(bytes), 256, 0, s>>>((uint32_t *)slot, (const uint32_t *)d_src, bytes);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return cuda_fail(e, "k_copy_small");
    memcpy(h_dst, slot, bytes);
    return FRB_OK;
}
inline int small_fill(void *d_dst, uint32_t word, size_t bytes, cudaStream_t s) {        // bytes and d_dst: multiples of 4
    if (bytes == 0) return FRB_OK;
    k_fill_small<<<small_grid(bytes), 256, 0, s>>>((uint32_t *)d_dst, word, bytes >> 2);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FRB_OK : cuda_fail(e, "k_fill_small");
}
#define FRB_TRY(call) do { int rc__ = (call); if (rc__ != FRB_OK) return rc__; } while (0)

// ---- optional kernel timing (bench roofline) ---------------------------------
struct ProfSlot { cudaEvent_t a = nullptr, b = nullptr; bool valid = false; };
static bool g_prof_on = false;
static ProfSlot g_prof[8];
inline void prof_begin(int which, cudaStream_t s) {
    if (!g_prof_on) return;
    ProfSlot &p = g_prof[which];
    if (!p.a) { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
    cudaEventRecord(p.a, s);
}
inline void prof_end(int which, cudaStream_t s) {
    if (!g_prof_on) return;
    cudaEventRecord(g_prof[which].b, s);
    g_prof[which].valid = true;
}

// ---- big-endian / bit helpers -------------------------------------------
__host__ __device__ __forceinline__ uint32_t bswap32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __byte_perm(v, 0, 0x0123);
#else
    return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
#endif
}

// CRC tables live in constant memory (CRC-8 poly 0x07, CRC-16 poly 0x8005; RFC 9639 9.1.8 / 9.3)
__constant__ uint8_t  c_crc8[256];
__constant__ uint16_t c_crc16[256];

// GF(2) helpers for CRC-16 combination: multiply a(x)*b(x) mod P, P = x^16+x^15+x^2+1
__host__ __device__ __forceinline__ uint32_t gf16_mul(uint32_t a, uint32_t b) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        r = (r << 1) ^ ((r & 0x8000u) ? 0x18005u : 0u);
        if (b & (0x8000u >> i)) r ^= a;
    }
    return r & 0xFFFFu;
}
// x^(8*nbytes) mod P
__host__ __device__ __forceinline__ uint32_t gf16_xpow8(uint64_t nbytes) {
    uint32_t result = 1;          // x^0
    uint32_t base = 0x0100u;      // x^8
    while (nbytes) {
        if (nbytes & 1) result = gf16_mul(result, base);
        base = gf16_mul(base, base);
        nbytes >>= 1;
    }
    return result;
}

// UTF-8-style coded number length (frame number field)
__host__ __device__ __forceinline__ int utf8_len(uint64_t v) {
    return v < 0x80 ? 1 : v < 0x800 ? 2 : v < 0x10000 ? 3 : v < 0x200000 ? 4 : v < 0x4000000 ? 5 : v < 0x80000000ull ? 6 : 7;
}

__host__ __device__ __forceinline__ uint32_t blocksize_code(uint32_t bs, int *hint_bytes) {
    *hint_bytes = 0;
    switch (bs) {
        case 192: return 1; case 576: return 2; case 1152: return 3; case 2304: return 4; case 4608: return 5;
        case 256: return 8; case 512: return 9; case 1024: return 10; case 2048: return 11; case 4096: return 12;
        case 8192: return 13; case 16384: return 14; case 32768: return 15;
    }
    if (bs <= 256) { *hint_bytes = 1; return 6; }
    *hint_bytes = 2; return 7;
}
__host__ __device__ __forceinline__ uint32_t samplerate_code(uint32_t sr, int *hint_kind) {
    *hint_kind = 0;
    switch (sr) {
        case 88200: return 1; case 176400: return 2; case 192000: return 3; case 8000: return 4;
        case 16000: return 5; case 22050: return 6; case 24000: return 7; case 32000: return 8;
        case 44100: return 9; case 48000: return 10; case 96000: return 11;
    }
    if (sr <= 255000 && sr % 1000 == 0) { *hint_kind = 1; return 12; }
    if (sr <= 655350 && sr % 10 == 0) { *hint_kind = 3; return 14; }
    if (sr <= 0xFFFF) { *hint_kind = 2; return 13; }
    return 0;
}
__host__ __device__ __forceinline__ uint32_t bps_code(uint32_t bps) {
    switch (bps) { case 8: return 1; case 12: return 2; case 16: return 4; case 20: return 5; case 24: return 6; case 32: return 7; }
    return 0;
}

}  // namespace frb
