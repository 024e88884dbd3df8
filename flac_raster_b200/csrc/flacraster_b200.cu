// flacraster_b200.cu -- unity build of libflacraster_b200.so (sm_100a only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <vector>
#include <cstdlib>
#include <cmath>
#include <mutex>
#include "frb_common.cuh"

namespace frb {
static int ensure_tables_impl();
}
#include "frb_normalize.cuh"
#include "frb_tilemap.cuh"
#include "frb_decode.cuh"
#include "frb_encode.cuh"
#include "frb_host.cuh"
#include "frb_hostparse.cuh"

namespace frb {
// The CRC-16 tables, built once on the host (no CUDA call: the host self-test uses them too).
static const CrcTables &host_crc_tables() {
    static CrcTables T;
    static std::once_flag once;
    std::call_once(once, [] {
        uint16_t t16[256];
        for (int i = 0; i < 256; i++) {
            uint16_t d = (uint16_t)(i << 8);
            for (int b = 0; b < 8; b++) d = (d & 0x8000) ? (uint16_t)((d << 1) ^ 0x8005) : (uint16_t)(d << 1);
            t16[i] = d;
        }
        for (int b = 0; b < 256; b++) {
            uint32_t c = t16[b];                                  // byte b followed by 0 zero bytes
            T.s4[b] = (uint16_t)c;
            for (int k = 1; k < 4; k++) { c = ((c << 8) & 0xFFFFu) ^ t16[c >> 8]; T.s4[k * 256 + b] = (uint16_t)c; }
        }
        for (int i = 0; i < 2048; i++) T.xp[i] = (uint16_t)gf16_xpow8((uint64_t)i);
    });
    return T;
}
// CRC tables are uploaded once per device.
static int ensure_tables_impl() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    FRB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return FRB_OK;
    uint8_t t8[256]; uint16_t t16[256];
    for (int i = 0; i < 256; i++) {
        uint8_t c = (uint8_t)i;
        for (int b = 0; b < 8; b++) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
        t8[i] = c;
        uint16_t d = (uint16_t)(i << 8);
        for (int b = 0; b < 8; b++) d = (d & 0x8000) ? (uint16_t)((d << 1) ^ 0x8005) : (uint16_t)(d << 1);
        t16[i] = d;
    }
    FRB_CUDA(cudaMemcpyToSymbol(c_crc8, t8, sizeof t8));
    FRB_CUDA(cudaMemcpyToSymbol(c_crc16, t16, sizeof t16));
    const CrcTables &T = host_crc_tables();
    FRB_CUDA(cudaMemcpyToSymbol(d_crct, &T, sizeof T));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return FRB_OK;
}
}  // namespace frb

extern "C" int frb_version(void) { return 1; }
extern "C" const char *frb_error_string(int status) {
    switch (status) {
        case FRB_OK: return "ok";
        case FRB_ERR_INVALID_ARG: return "invalid argument";
        case FRB_ERR_CUDA: return "CUDA error";
        case FRB_ERR_UNSUPPORTED: return "unsupported stream or parameter";
        case FRB_ERR_BAD_STREAM: return "malformed FLAC stream";
        case FRB_ERR_CRC: return "CRC mismatch";
        case FRB_ERR_OVERFLOW: return "output buffer too small";
        case FRB_ERR_NO_DEVICE: return "no CUDA device";
    }
    return "unknown";
}
extern "C" const char *frb_last_cuda_error(void) { return frb::g_last_cuda_error; }
extern "C" int frb_device_count(int *count) {
    if (!count) return FRB_ERR_INVALID_ARG;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; cudaGetLastError(); return FRB_ERR_NO_DEVICE; }
    return FRB_OK;
}
// Host self-test of the table-free CRC-16 (frb_crc16.cuh): runs the per-lane code of warp_crc16 for lanes 0..31 on the host, for
// bytes [a, e) of `bytes` (readable up to the next 16-byte boundary after e, base 16-byte aligned).  No GPU needed.
extern "C" int frb_selftest_crc16(const uint8_t *bytes, uint64_t a, uint64_t e, int group, uint32_t *crc) {
    if (!bytes || !crc || e < a || (reinterpret_cast<uintptr_t>(bytes) & 15u) || e - a >= (1ull << 32) || (group != 32 && group != 128))
        return FRB_ERR_INVALID_ARG;
    const frb::CrcTables &T = frb::host_crc_tables();
    uint32_t v = 0;
    for (int lane = 0; lane < group; lane++)
        v ^= group == 32 ? frb::lane_crc16<30>(bytes, a, e, T.s4, T.xp, lane) : frb::lane_crc16<120>(bytes, a, e, T.s4, T.xp, lane);
    *crc = v;
    return FRB_OK;
}
extern "C" uint64_t frb_launch_count(void) { return frb::g_launches.load(); }

// Small transfers through pinned staging + a copy kernel (see frb_common.cuh: they must not queue behind bulk DMA).
extern "C" int frb_small_upload(void *d_dst, const void *h_src, size_t bytes, void *stream) {
    if (!d_dst || !h_src) return FRB_ERR_INVALID_ARG;
    return frb::small_upload(d_dst, h_src, bytes, (cudaStream_t)stream);
}
extern "C" int frb_small_download(void *h_dst, const void *d_src, size_t bytes, void *stream) {
    if (!h_dst || !d_src) return FRB_ERR_INVALID_ARG;
    return frb::small_download(h_dst, d_src, bytes, (cudaStream_t)stream);
}

extern "C" int frb_profile_enable(int on) { frb::g_prof_on = on != 0; return FRB_OK; }
extern "C" int frb_profile_last_ms(int which, float *ms) {
    if (which < 0 || which > 7 || !ms) return FRB_ERR_INVALID_ARG;
    frb::ProfSlot &p = frb::g_prof[which];
    if (!p.valid) return FRB_ERR_INVALID_ARG;
    FRB_CUDA(cudaEventSynchronize(p.b));
    FRB_CUDA(cudaEventElapsedTime(ms, p.a, p.b));
    return FRB_OK;
}

namespace frb {
__global__ void __launch_bounds__(128) k_debug_spin(unsigned long long ns) {
    extern __shared__ uint8_t spin_smem[];
    if (threadIdx.x == 0) spin_smem[0] = 0;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { __nanosleep(1000); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}
}  // namespace frb
extern "C" int frb_debug_spin(uint32_t ctas, uint32_t smem_bytes, uint64_t ns, void *stream) {
    using namespace frb;
    if (!ctas || smem_bytes > 200 * 1024 || ns > 2000000000ull) return FRB_ERR_INVALID_ARG;
    FRB_CUDA(cudaFuncSetAttribute(k_debug_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_debug_spin<<<ctas, 128, smem_bytes, (cudaStream_t)stream>>>(ns);
    FRB_LAUNCH_CHECK("k_debug_spin");
    return FRB_OK;
}
