// frb_crc16.cuh -- included inside namespace frb (after FrameLoc/locate_frame are defined).
//
// CRC-16 (poly 0x8005, init 0, MSB first).  CRC(M) = M(x) * x^16 mod P, so
// CRC(A||B) = CRC(A) * x^(8|B|) + CRC(B) and leading zero bytes are neutral.  A warp owns a byte
// range; lane l < 30 folds 16-byte chunks l, l+30, ... without tables (CrcFold below: a row of 30
// chunks is 3840 bits and x^3840 = x^256 + 1 modulo the degree-15 factor of P).  Lane results are
// weighted by x^(128*(29-l)); the ragged last row and the tail bytes by small tabulated powers
// (slice-by-4 tables, once per frame).
struct CrcTables {
    uint16_t s4[4 * 256];    // slice-by-4: s4[k*256+b] = CRC of byte b followed by k zero bytes
    uint16_t xp[2048];       // xp[i] = x^(8*i) mod P
};
// The tables live in global memory: every CTA copies them to shared memory with coalesced 16-byte loads (L2
// hits).  They used to sit in __constant__ memory, where the copy's per-thread addresses serialise in the
// constant cache (32 passes per warp instruction): ~1 us per CTA, which dominated k_emit_frames for
// single-channel frames (7.4 ms for 262 144 frames of C5, profiles/r01_bench_c5_v1.json).
__device__ __align__(16) CrcTables d_crct;

__host__ __device__ __forceinline__ uint32_t crc16_word(uint32_t crc, uint32_t w, const uint16_t *s4) {
    const uint32_t x = w ^ (crc << 16);
    return (uint32_t)s4[3 * 256 + (x >> 24)] ^ s4[2 * 256 + ((x >> 16) & 0xFF)] ^ s4[256 + ((x >> 8) & 0xFF)] ^ s4[x & 0xFF];
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
__host__ __device__ __forceinline__ void crc_mask_head(uint32_t (&w)[4], uint32_t head) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int drop = (int)head - 4 * q;               // leading bytes of word q to zero
        if (drop >= 4) w[q] = 0; else if (drop > 0) w[q] &= 0xFFFFFFFFu >> (8 * drop);
    }
}

// ---- table-free CRC accumulation ("fold" form) ------------------------------------------------------------------------------
// P = x^16 + x^15 + x^2 + 1 = (x + 1) * Q with Q = x^15 + x + 1.  Modulo Q, x^15 = x + 1, and squaring is linear over GF(2), so
//     x^(15 * 2^j) = (x + 1)^(2^j) = x^(2^j) + 1        (x^240 = x^16 + 1, x^480 = x^32 + 1, x^3840 = x^256 + 1, ...):
// multiplying by these powers is two word-aligned shifts and an XOR -- no table, no carry-less multiply.  A lane that folds the
// 16-byte chunks l, l+30, l+60, ... (rows of 30 chunks = 3840 bits) keeps S = sum of its chunks * x^(3840 * rows to go) as an
// UNREDUCED 480-bit value (15 words): one step is S = S * (x^256 + 1) + chunk with the words above 480 bits brought back by
// x^480 = x^32 + 1 -- 18 three-input XORs per chunk against 16 shared-memory table lookups (2.5-way bank conflicts on random
// indices: k_crc16_frames and k_emit_frames were bound by them) and ~50 ALU instructions of the slice-by-4 update.  Modulo x + 1
// a polynomial is its parity, whatever the positions.  At the end of a frame S is folded to 15 bits (x^240, x^120, x^60, x^30,
// x^30, x^15, x^15) and the two residues give M mod P by the Chinese remainder theorem: r_Q, plus Q when the parities differ.
// Plain C so that the host self-test (frb_selftest_crc16) runs the same code.
struct CrcFold {
    uint32_t s[15];      // s[0] = least significant word
    uint32_t par;        // XOR of every word folded in (its parity = M mod (x + 1))
};
__host__ __device__ __forceinline__ void crcfold_init(CrcFold &f) {
#pragma unroll
    for (int i = 0; i < 15; i++) f.s[i] = 0;
    f.par = 0;
}
// S = S * x^3840 + chunk (mod Q); w[0] holds the chunk's first four bytes (big-endian words)
__host__ __device__ __forceinline__ void crcfold_step30(CrcFold &f, const uint32_t (&w)[4]) {
    uint32_t n[15];
    const uint32_t *s = f.s;
    // word i of S*x^256 + S: s[i] ^ s[i-8]; words 15..22 (= s[7..14]) come back as x^(32 (j+1)) + x^(32 j)
    n[0] = s[0] ^ s[7] ^ w[3];
    n[1] = s[1] ^ s[8] ^ s[7] ^ w[2];
    n[2] = s[2] ^ s[9] ^ s[8] ^ w[1];
    n[3] = s[3] ^ s[10] ^ s[9] ^ w[0];
    n[4] = s[4] ^ s[11] ^ s[10];
    n[5] = s[5] ^ s[12] ^ s[11];
    n[6] = s[6] ^ s[13] ^ s[12];
    n[7] = s[7] ^ s[14] ^ s[13];
    n[8] = s[8] ^ s[0] ^ s[14];
#pragma unroll
    for (int i = 9; i < 15; i++) n[i] = s[i] ^ s[i - 8];
    f.par ^= w[0] ^ w[1] ^ w[2] ^ w[3];
#pragma unroll
    for (int i = 0; i < 15; i++) f.s[i] = n[i];
}
// S = S * x^15360 + chunk (mod Q): rows of 120 chunks (a 128-thread group per frame).  x^15360 = x^1024 + 1; word i of S*x^1024 sits
// at word i+32 and comes back by x^960 = x^64 + 1 to words i+4 and i+2, those past word 14 once more by x^480 = x^32 + 1.
__host__ __device__ __forceinline__ void crcfold_step120(CrcFold &f, const uint32_t (&w)[4]) {
    uint32_t n[15];
    const uint32_t *s = f.s;
    n[0] = s[0] ^ s[11] ^ s[13] ^ w[3];
    n[1] = s[1] ^ s[11] ^ s[12] ^ s[13] ^ s[14] ^ w[2];
    n[2] = s[2] ^ s[0] ^ s[12] ^ s[13] ^ s[14] ^ w[1];
    n[3] = s[3] ^ s[1] ^ s[13] ^ s[14] ^ w[0];
    n[4] = s[4] ^ s[2] ^ s[0] ^ s[14];
#pragma unroll
    for (int i = 5; i < 15; i++) n[i] = s[i] ^ s[i - 2] ^ s[i - 4];
    f.par ^= w[0] ^ w[1] ^ w[2] ^ w[3];
#pragma unroll
    for (int i = 0; i < 15; i++) f.s[i] = n[i];
}
template <int ROW>
__host__ __device__ __forceinline__ void crcfold_step(CrcFold &f, const uint32_t (&w)[4]) {
    static_assert(ROW == 30 || ROW == 120, "row lengths with a two-term multiplier");
    if (ROW == 30) crcfold_step30(f, w); else crcfold_step120(f, w);
}
// r = lo(K bits) + hi + hi * x^J for r of NW words (x^K = x^J + 1 mod Q); K, J compile-time
template <int NW, int K, int J>
__host__ __device__ __forceinline__ void crcfold_halve(uint32_t (&r)[NW]) {
    uint32_t hi[NW], out[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) {                        // hi = r >> K
        const int lo_w = i + K / 32, sh = K % 32;
        const uint32_t a = lo_w < NW ? r[lo_w] : 0u, b = lo_w + 1 < NW ? r[lo_w + 1] : 0u;
        hi[i] = sh ? (a >> sh) | (b << (32 - sh)) : a;
    }
#pragma unroll
    for (int i = 0; i < NW; i++) {                        // lo ^ hi ^ (hi << J)
        const uint32_t lo = i < K / 32 ? r[i] : i == K / 32 ? (K % 32 ? r[i] & ((1u << (K % 32)) - 1u) : 0u) : 0u;
        const int src = i - J / 32, sh = J % 32;
        const uint32_t a = src >= 0 && src < NW ? hi[src] : 0u, b = src - 1 >= 0 && src - 1 < NW ? hi[src - 1] : 0u;
        out[i] = lo ^ hi[i] ^ (sh ? (a << sh) | (b >> (32 - sh)) : a);
    }
#pragma unroll
    for (int i = 0; i < NW; i++) r[i] = out[i];
}
// M mod P (16 bits) of what was folded in; the CRC register after those bytes is this times x^16
__host__ __device__ __forceinline__ uint32_t crcfold_finish(const CrcFold &f) {
    uint32_t a[15];
#pragma unroll
    for (int i = 0; i < 15; i++) a[i] = f.s[i];
    crcfold_halve<15, 240, 16>(a);                        // 480 -> 256 bits
    uint32_t b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = a[i];
    crcfold_halve<8, 120, 8>(b);                          // -> 144
    uint32_t c[5];
#pragma unroll
    for (int i = 0; i < 5; i++) c[i] = b[i];
    crcfold_halve<5, 60, 4>(c);                           // -> 88
    uint32_t d[3] = {c[0], c[1], c[2]};
    crcfold_halve<3, 30, 2>(d);                           // -> 60
    uint32_t e[2] = {d[0], d[1]};
    crcfold_halve<2, 30, 2>(e);                           // -> 32
    uint32_t r = e[0];
    uint32_t h = r >> 15;
    r = (r & 0x7FFFu) ^ h ^ (h << 1);                     // -> 18
    h = r >> 15;
    r = (r & 0x7FFFu) ^ h ^ (h << 1);                     // -> 15 bits: M mod Q
    uint32_t p = f.par ^ r;                               // parity(M) ^ parity(r)
    p ^= p >> 16; p ^= p >> 8; p ^= p >> 4; p ^= p >> 2; p ^= p >> 1;
    return r ^ ((p & 1u) ? 0x8003u : 0u);                 // + Q: the only other value that is r mod Q; Q has odd parity
}

// One lane's share of the CRC-16 of bytes [a, e) of `bytes` (16-byte aligned base): the XOR over lanes 0..31 is the CRC.
// The buffer must be readable up to the 16-byte boundary after e.
__host__ __device__ __forceinline__ uint4 crc_ld_chunk(const uint4 *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
// ROW = 30: a warp per range (lanes 30, 31 only help with the tail); ROW = 120: a group of 128 threads.
template <int ROW>
__host__ __device__ __forceinline__ uint32_t lane_crc16(const uint8_t *__restrict__ bytes, uint64_t a, uint64_t e,
                                                        const uint16_t *s4, const uint16_t *xp, int lane) {
    constexpr int kLast = ROW == 30 ? 31 : 127;
    const uint64_t a0 = a & ~(uint64_t)15;
    const uint32_t nfull = (uint32_t)((e - a0) >> 4);     // full 16-byte chunks from a0 (first one front-masked); frames are < 2^32 bytes
    const uint32_t tl = (uint32_t)((e - a0) & 15);        // tail bytes after the last full chunk
    const uint32_t rows = nfull / (uint32_t)ROW, rem = nfull - rows * (uint32_t)ROW;
    const uint4 *chunks = reinterpret_cast<const uint4 *>(bytes + a0);
    const uint32_t head = (uint32_t)(a - a0);
    uint32_t v = 0;
    if (lane < ROW) {
        CrcFold F;
        crcfold_init(F);
        for (uint32_t m = 0; m < rows; m++) {
            const uint32_t idx = m * (uint32_t)ROW + (uint32_t)lane;
            const uint4 q4 = crc_ld_chunk(chunks + idx);
            uint32_t w[4] = {bswap32(q4.x), bswap32(q4.y), bswap32(q4.z), bswap32(q4.w)};
            if (idx == 0 && head) crc_mask_head(w, head);
            crcfold_step<ROW>(F, w);
        }
        // x^16 (message -> CRC register), the lane's weight inside a row, then everything after the full rows: 16*rem + tl bytes
        if (rows) v = gf16_mul(gf16_mul(crcfold_finish(F), xp[16 * (ROW - 1 - lane) + 2]), xp[16 * rem + tl]);
        if ((uint32_t)lane < rem) {
            const uint32_t idx = rows * (uint32_t)ROW + (uint32_t)lane;
            const uint4 q4 = crc_ld_chunk(chunks + idx);
            uint32_t w[4] = {bswap32(q4.x), bswap32(q4.y), bswap32(q4.z), bswap32(q4.w)};
            if (idx == 0 && head) crc_mask_head(w, head);
            uint32_t c = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) c = crc16_word(c, w[q], s4);
            v ^= gf16_mul(c, xp[16 * (rem - 1 - lane) + tl]);
        }
    }
    if (lane == kLast && tl) {
        // tail bytes (weight 1); when the whole range is shorter than one chunk the head mask applies too
        const uint4 q4 = crc_ld_chunk(chunks + nfull);
        const uint32_t w[4] = {bswap32(q4.x), bswap32(q4.y), bswap32(q4.z), bswap32(q4.w)};
        uint32_t c = 0;
        for (uint32_t q = 0; q < tl; q++) {
            uint32_t byte = (w[q >> 2] >> (24 - 8 * (q & 3))) & 0xFF;
            if (nfull == 0 && q < head) byte = 0;
            c = ((c << 8) & 0xFFFFu) ^ s4[((c >> 8) ^ byte) & 0xFF];
        }
        v ^= c;
    }
    return v;
}
// CRC-16 of bytes [a, e), computed by one warp; every lane returns it.
__device__ __forceinline__ uint32_t warp_crc16(const uint8_t *__restrict__ bytes, uint64_t a, uint64_t e,
                                               const CrcTables *T, int lane) {
    uint32_t v = lane_crc16<30>(bytes, a, e, T->s4, T->xp, lane);
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
