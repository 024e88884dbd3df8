// int_peak.cu -- measured integer issue rates of one B200 (the issue-rate roofline of the bit-parsing
// kernels needs a measured INT32 peak, not a data-sheet guess).  Each kernel keeps 8 independent dependency chains per
// thread so that the pipes, not latencies, limit it.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kChains = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    uint32_t a[kChains], b = seed | 1u, c = seed * 2654435761u + 12345u;
#pragma unroll
    for (int i = 0; i < kChains; i++) a[i] = threadIdx.x * 31u + i * 7u + seed;
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < kChains; i++) {
            if (MODE == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));                       // IADD3 (ALU pipe)
            else if (MODE == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); // LOP3 (ALU pipe)
            else if (MODE == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));     // IMAD (FMA pipe)
            else if (MODE == 3) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b)); // SHF (ALU pipe)
            else if (MODE == 4) {                                                                              // IADD3 + IMAD alternating
                if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            } else if (MODE == 5) asm volatile("bfind.u32 %0, %0;" : "+r"(a[i]));                              // FLO (XU pipe)
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < kChains; i++) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
static void run(const char *name, int sms, int clock_khz) {
    const int blocks = sms * 8, threads = 256;
    uint32_t *out;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 3u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(out, 3u + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double lane_ops = (double)blocks * threads * kIters * kChains;
    const double tops = lane_ops / (best * 1e-3) / 1e12;
    const double warp_inst_per_clk_per_sm = lane_ops / 32.0 / (best * 1e-3) / ((double)clock_khz * 1e3) / sms;
    printf("%-28s %8.3f ms  %7.2f Tops/s (lane ops)  %5.2f warp-inst/clk/SM at %d MHz\n", name, best, tops, warp_inst_per_clk_per_sm, clock_khz / 1000);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s: %d SMs, %d kHz\n", p.name, p.multiProcessorCount, clk);
    run<0>("IADD3 (add.u32)", p.multiProcessorCount, clk);
    run<1>("LOP3", p.multiProcessorCount, clk);
    run<3>("SHF (funnel shift)", p.multiProcessorCount, clk);
    run<2>("IMAD (mad.lo.u32)", p.multiProcessorCount, clk);
    run<4>("IADD3 + IMAD alternating", p.multiProcessorCount, clk);
    run<5>("FLO (bfind.u32)", p.multiProcessorCount, clk);
    return 0;
}
