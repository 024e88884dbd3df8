// frb_crc16.cuh -- included inside namespace frb (after FrameLoc/locate_frame are defined).
//
// CRC-16 (poly 0x8005, init 0, MSB first).  CRC(M) = M(x) * x^16 mod P, so
// CRC(A||B) = CRC(A) * x^(8|B|) + CRC(B) and leading zero bytes are neutral.  A warp owns a byte
// range; lane l folds 16-byte chunks l, l+32, ... with a Horner step acc = acc * x^(8*496) followed
// by four slice-by-4 word updates (together x^(8*512) per row of 32 chunks).  Lane results are
// weighted by x^(128*(31-l)); the ragged last row and the tail bytes by small tabulated powers.
struct CrcTables {
    uint16_t s4[4 * 256];    // slice-by-4: s4[k*256+b] = CRC of byte b followed by k zero bytes
    uint16_t k496[512];      // multiply by x^(8*496): [h] for the high byte, [256+l] for the low byte
    uint16_t k2032[512];     // multiply by x^(8*2032) (128-thread CTAs: 128 chunks per row)
    uint16_t xp[2048];       // xp[i] = x^(8*i) mod P
};
// The tables live in global memory: every CTA copies them to shared memory with coalesced 16-byte loads (L2
// hits).  They used to sit in __constant__ memory, where the copy's per-thread addresses serialise in the
// constant cache (32 passes per warp instruction): ~1 us per CTA, which dominated k_emit_frames for
// single-channel frames (7.4 ms for 262 144 frames of C5, profiles/r01_bench_c5_v1.json).
__device__ __align__(16) CrcTables d_crct;

__device__ __forceinline__ uint32_t crc16_word(uint32_t crc, uint32_t w, const uint16_t *s4) {
    const uint32_t x = w ^ (crc << 16);
    return (uint32_t)s4[3 * 256 + (x >> 24)] ^ s4[2 * 256 + ((x >> 16) & 0xFF)] ^ s4[256 + ((x >> 8) & 0xFF)] ^ s4[x & 0xFF];
}
__device__ __forceinline__ uint32_t crc16_mul496(uint32_t acc, const uint16_t *k496) {
    return (uint32_t)k496[acc >> 8] ^ k496[256 + (acc & 0xFF)];
}
__device__ __forceinline__ uint32_t crc16_words4(uint32_t acc, const uint32_t (&w)[4], const uint16_t *s4) {
#pragma unroll
    for (int q = 0; q < 4; q++) acc = crc16_word(acc, w[q], s4);
    return acc;
}
__device__ __forceinline__ void crc_tables_to_smem(CrcTables *dst) {
    static_assert(sizeof(CrcTables) % 16 == 0, "vector copy");
    const uint4 *s = reinterpret_cast<const uint4 *>(&d_crct);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (uint32_t i = threadIdx.x; i < sizeof(CrcTables) / 16; i += blockDim.x) d[i] = __ldg(s + i);
}
__device__ __forceinline__ void crc_mask_head(uint32_t (&w)[4], uint32_t head) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int drop = (int)head - 4 * q;               // leading bytes of word q to zero
        if (drop >= 4) w[q] = 0; else if (drop > 0) w[q] &= 0xFFFFFFFFu >> (8 * drop);
    }
}

// CRC-16 of bytes [a, e) of `bytes` (16-byte aligned base), computed by one warp; every lane returns it.
// The buffer must be readable up to the 16-byte boundary after e.
__device__ __forceinline__ uint32_t warp_crc16(const uint8_t *__restrict__ bytes, uint64_t a, uint64_t e,
                                               const CrcTables *T, int lane) {
    const uint64_t a0 = a & ~(uint64_t)15;
    const uint64_t nfull = (e - a0) >> 4;                 // full 16-byte chunks from a0 (first one front-masked)
    const uint32_t tl = (uint32_t)((e - a0) & 15);        // tail bytes after the last full chunk
    const uint64_t rows = nfull >> 5;
    const uint32_t rem = (uint32_t)(nfull & 31);
    const uint4 *chunks = reinterpret_cast<const uint4 *>(bytes + a0);
    const uint32_t head = (uint32_t)(a - a0);
    uint32_t acc = 0;
    for (uint64_t m = 0; m < rows; m++) {
        const uint64_t idx = m * 32 + lane;
        const uint4 v = __ldg(chunks + idx);
        uint32_t w[4] = {bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w)};
        if (idx == 0 && head) crc_mask_head(w, head);
        acc = crc16_mul496(acc, T->k496);
#pragma unroll
        for (int q = 0; q < 4; q++) acc = crc16_word(acc, w[q], T->s4);
    }
    // weight of lane l inside a row, then everything after the full rows: 16*rem + tl bytes
    uint32_t v = gf16_mul(gf16_mul(acc, T->xp[16 * (31 - lane)]), T->xp[16 * rem + tl]);
    if ((uint32_t)lane < rem) {
        const uint64_t idx = rows * 32 + lane;
        const uint4 q4 = __ldg(chunks + idx);
        uint32_t w[4] = {bswap32(q4.x), bswap32(q4.y), bswap32(q4.z), bswap32(q4.w)};
        if (idx == 0 && head) crc_mask_head(w, head);
        uint32_t c = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) c = crc16_word(c, w[q], T->s4);
        v ^= gf16_mul(c, T->xp[16 * (rem - 1 - lane) + tl]);
    }
    if (lane == 31 && tl) {
        // tail bytes (weight 1); when the whole range is shorter than one chunk the head mask applies too
        const uint4 q4 = __ldg(chunks + nfull);
        const uint32_t w[4] = {bswap32(q4.x), bswap32(q4.y), bswap32(q4.z), bswap32(q4.w)};
        uint32_t c = 0;
        for (uint32_t q = 0; q < tl; q++) {
            uint32_t byte = (w[q >> 2] >> (24 - 8 * (q & 3))) & 0xFF;
            if (nfull == 0 && q < head) byte = 0;
            c = ((c << 8) & 0xFFFFu) ^ T->s4[((c >> 8) ^ byte) & 0xFF];
        }
        v ^= c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
k_crc16_frames(const uint8_t *__restrict__ bytes, const DecStreamDev *__restrict__ streams, uint32_t n_streams,
               uint32_t blocksize, uint32_t total_frames, const unsigned long long *__restrict__ frame_pos,
               uint32_t *__restrict__ status) {
    __shared__ __align__(16) CrcTables T;
    crc_tables_to_smem(&T);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < total_frames; f += warps) {    // warp-uniform
        DecStreamDev st;
        const FrameLoc L = locate_frame(streams, n_streams, blocksize, f, frame_pos, &st);
        if (!L.ok) continue;                              // counted as missing by the decoder
        const uint32_t crc = warp_crc16(bytes, L.start, L.end - 2, &T, lane);
        if (lane == 0) {
            const uint32_t want = ((uint32_t)bytes[L.end - 2] << 8) | bytes[L.end - 1];
            if (crc != want) atomicAdd(&status[1], 1u);
        }
    }
}
