// flacraster_b200.cu -- unity build of libflacraster_b200.so (sm_100a only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <vector>
#include <cmath>
#include <mutex>
#include "frb_common.cuh"

namespace frb {
static int ensure_tables_impl();
}
#include "frb_normalize.cuh"
#include "frb_decode.cuh"
#include "frb_encode.cuh"
#include "frb_host.cuh"

namespace frb {
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
extern "C" uint64_t frb_launch_count(void) { return frb::g_launches.load(); }

extern "C" int frb_profile_enable(int on) { frb::g_prof_on = on != 0; return FRB_OK; }
extern "C" int frb_profile_last_ms(int which, float *ms) {
    if (which < 0 || which > 3 || !ms) return FRB_ERR_INVALID_ARG;
    frb::ProfSlot &p = frb::g_prof[which];
    if (!p.valid) return FRB_ERR_INVALID_ARG;
    FRB_CUDA(cudaEventSynchronize(p.b));
    FRB_CUDA(cudaEventElapsedTime(ms, p.a, p.b));
    return FRB_OK;
}
