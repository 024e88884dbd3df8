// Dependent-chain latency of the integer ops the Rice parser is made of (one warp, 4096 dependent ops, clock64).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat.bin lat.cu && ./lat.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CHAIN(name, body)                                                                       \
    __global__ void k_##name(uint32_t *out, long long *cyc, uint32_t seed, uint32_t b) {        \
        uint32_t x = seed + threadIdx.x, y = b;                                                 \
        __shared__ uint32_t sm[1024];                                                           \
        for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i * 7 + 3) & 1023;                \
        __syncwarp();                                                                           \
        long long t0 = clock64();                                                               \
        _Pragma("unroll 16") for (int i = 0; i < 4096; i++) { body; }                           \
        long long t1 = clock64();                                                               \
        out[threadIdx.x] = x + y; if (threadIdx.x == 0) *cyc = t1 - t0;                         \
    }
CHAIN(iadd, asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y)))
CHAIN(lop3, asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(y)))
CHAIN(shf, asm volatile("shf.l.wrap.b32 %0, %0, %1, %1;" : "+r"(x) : "r"(y)))
CHAIN(imad, asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(y)))
CHAIN(bfind, asm volatile("bfind.u32 %0, %0; or.b32 %0, %0, %1;" : "+r"(x) : "r"(y)))          // bfind + 1 lop
CHAIN(clz, asm volatile("clz.b32 %0, %0; or.b32 %0, %0, %1;" : "+r"(x) : "r"(y)))
CHAIN(popc, asm volatile("popc.b32 %0, %0; or.b32 %0, %0, %1;" : "+r"(x) : "r"(y)))
CHAIN(brev, asm volatile("brev.b32 %0, %0;" : "+r"(x)))
CHAIN(cvtf, { float f; asm volatile("cvt.rz.f32.u32 %0, %1;" : "=f"(f) : "r"(x)); x = __float_as_uint(f) | y; })   // I2F + lop
CHAIN(prmt, asm volatile("prmt.b32 %0, %0, %1, 0x0123;" : "+r"(x) : "r"(y)))
CHAIN(sel, asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %0, %1, p;}" : "+r"(x) : "r"(y)))   // setp + selp
CHAIN(lds, x = sm[x & 1023])
CHAIN(shfl, x = __shfl_xor_sync(0xFFFFFFFFu, x, 1))
CHAIN(vimnmx, asm volatile("max.u32 %0, %0, %1;" : "+r"(x) : "r"(y)))
CHAIN(bfe, asm volatile("bfe.s32 %0, %0, 0, 9;" : "+r"(x)))
int main() {
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
#define RUN(name, extra) k_##name<<<1, 32>>>(out, cyc, 12345u, 3u); k_##name<<<1, 32>>>(out, cyc, 12345u, 3u); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-8s %7.2f cycles per step%s\n", #name, h / 4096.0, extra);
    RUN(iadd, "") RUN(lop3, "") RUN(shf, "") RUN(imad, "") RUN(vimnmx, "") RUN(prmt, "") RUN(bfe, "") RUN(brev, "")
    RUN(bfind, "  (bfind + or)") RUN(clz, "  (clz + or)") RUN(popc, "  (popc + or)") RUN(cvtf, "  (cvt.rz.f32.u32 + or)") RUN(sel, "  (setp + selp)")
    RUN(lds, "") RUN(shfl, "")
    return 0;
}
