// frb_decode_kernels.cuh -- included inside namespace frb by frb_decode.cuh.
//
// v2 decode pipeline (justified by profiles/r01_ncu_dec_v1_raw_subset.csv: the v1 thread-per-frame
// kernel ran at 9.7 % occupancy on 8-band tiles and spent 228 instructions per sample):
//   k_skim_subframes    (channels > 1 only) one thread per frame walks the Rice codes without
//                       reconstructing anything and records each subframe's bit offset
//   k_decode_subframes  one thread per SUBFRAME; 32-bit funnel-shift bit window with a prefetched
//                       third word; LPC history in registers (sliding window, 4 samples per group,
//                       taps padded to the warp's order class 4/8/12); 128-bit stores
//   k_crc16_frames      one warp per frame, 16-byte chunks per lane, slice-by-4 tables in shared
//                       memory, Horner combination with x^(8*512) and a final x^(8*n) weight

// ---- cp.async / shared-memory helpers ------------------------------------------------------------------
constexpr int kDecThreads = 128;
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t smem_addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(smem_addr) : "memory");
    return v;
}

__device__ __forceinline__ int32_t unzigzag(uint32_t u) { return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u); }

struct FrameLoc {
    uint64_t start, end;      // byte range [start, end) of the frame incl. CRC-16
    uint32_t k, n, stream;    // frame number in stream, blocksize of this frame
    bool ok;
};

__device__ __forceinline__ FrameLoc locate_frame(const DecStreamDev *__restrict__ streams, uint32_t n_streams,
                                                 uint32_t blocksize, uint32_t f,
                                                 const unsigned long long *__restrict__ frame_pos, DecStreamDev *st_out) {
    uint32_t lo = 0, hi = n_streams - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (streams[mid].frame_base <= f) lo = mid; else hi = mid - 1;
    }
    const DecStreamDev st = streams[lo];
    FrameLoc L;
    L.stream = lo;
    L.k = f - st.frame_base;
    const unsigned long long p0 = frame_pos[f];
    const unsigned long long p1 = (L.k + 1 < st.n_frames) ? frame_pos[f + 1] : (st.byte_offset + st.byte_length);
    L.n = (L.k + 1 < st.n_frames) ? blocksize : (uint32_t)(st.n_samples - (uint64_t)L.k * blocksize);
    L.ok = !(p0 == kNoPos || p1 == kNoPos || p1 <= p0 + 4);
    L.start = p0; L.end = p1;
    *st_out = st;
    return L;
}

// status words: 0 frames_missing, 1 crc16_errors, 2 parse_errors, 3 frames_decoded, 4 order_overflow

// ---- bit reader + skim: subframe bit offsets for multi-channel frames ------------------------------
// sub_bitoff[f*channels + c] = bit offset of subframe c from the frame start; 0 marks a bad frame.
// One thread per frame walks subframes 0..C-2 (the last one only needs its start; its end is checked by
// the decode kernel).
//
// With 32 independent frames per warp every data-dependent branch is taken by SOME lane in almost every
// step, so the warp pays for the union of all paths: the v3 skim (shared-memory word ring, one Rice code
// per loop trip, refill branch) and a v4 trial (16-byte ring slots + register queue) executed 50-100 warp
// instructions per code (5.8 / 12.8 ms on C3, profiles/r01_ncu_skim_v4.txt).  This version keeps the hot
// loop free of data-dependent branches:
//   * codes are skipped in batches of kSkimBatch with predication (lane inactive once its partition is
//     exhausted); a code longer than 32 bits is the only branch in the batch;
//   * the window is (hi, lo) plus a pre-loaded next word; crossing a word boundary is three selects and a
//     predicated 4-byte LDS from the thread's ring, never a wait;
//   * the ring (16-byte chunks, cp.async) is topped up ONCE per batch at a warp-uniform point with
//     predicated copies, and one wait_group per batch covers every word the batch can touch.
constexpr int kSkimRing = 16;                                   // 16-byte chunks per thread (256 contiguous bytes)
constexpr int kSkimBatch = 8;                                   // codes per batch: <= 8 words = 2 chunks

// 64-bit left-aligned window (hi:lo, the top vb >= 32 bits valid) + two pre-loaded words.  The dependent chain
// per code is clz -> add -> funnel shift (-> or, when a word is merged); everything else hangs off it.
// Ring: the thread's 16 chunks are contiguous in shared memory, chunk index XOR-swizzled with the lane so that
// lanes reading the same word offset spread over the banks.
struct BitReader {
    const uint4 *gq;         // 16-byte view of the input buffer (16-byte aligned base)
    uint32_t sbase;          // shared-space address of this thread's 256-byte ring
    uint32_t swz;            // (lane & 15) << 4
    uint32_t wnext;          // word index (32-bit words from the buffer base) of the next word to load; nx = wnext-2, nx2 = wnext-1
    uint32_t cissue;         // next chunk to copy into the ring
    uint32_t qlast;          // copies are clamped to this chunk (16 readable bytes follow the frame end)
    uint32_t hi, lo, nx, nx2raw;   // nx2raw: word wnext-1 as loaded (little-endian), byte-swapped when it becomes nx
    int32_t vb;              // valid bits in hi:lo
    __device__ __forceinline__ uint32_t load_raw(uint32_t a) const { return lds_u32(sbase + ((((a << 2) & 252u)) ^ swz)); }
    __device__ __forceinline__ uint32_t load_word(uint32_t a) const { return bswap32(load_raw(a)); }
    __device__ __forceinline__ void copy_chunk(uint32_t c) const {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sbase + (((c & (kSkimRing - 1)) << 4) ^ swz)),
                     "l"(gq + min(c, qlast)) : "memory");
    }
    // Copy ahead without overwriting anything still needed: chunk c may replace chunk c - kSkimRing once that one
    // lies before the chunk of the oldest pre-loaded word.  At most two chunks per call (a batch consumes at most two).
    __device__ __forceinline__ void top_up() {
        const uint32_t lim = ((wnext - 2) >> 2) + kSkimRing - 1;
#pragma unroll
        for (int r = 0; r < 2; r++)
            if (cissue < lim) { copy_chunk(cissue); cissue++; }
        cp_async_commit();
        // a chunk is read at least (kSkimRing - 3) / 2 batches after it was requested
        cp_async_wait<(kSkimRing - 3) / 2 - 1>();
    }
    __device__ __forceinline__ void init(uint32_t ring_saddr, uint32_t lane, const uint8_t *base, uint64_t bitpos, uint64_t byte_end) {
        gq = (const uint4 *)base;
        sbase = ring_saddr;
        swz = (lane & 15u) << 4;
        const uint32_t w = (uint32_t)(bitpos >> 5);
        qlast = (uint32_t)(byte_end >> 4);
        cissue = w >> 2;
        for (int j = 0; j < kSkimRing - 1; j++) { copy_chunk(cissue); cissue++; }
        cp_async_commit();
        cp_async_wait<0>();
        hi = load_word(w); lo = load_word(w + 1); nx = load_word(w + 2); nx2raw = load_raw(w + 3);
        wnext = w + 4;
        vb = 64;
        consume((uint32_t)bitpos & 31u);
    }
    __device__ __forceinline__ uint64_t bitpos() const { return (uint64_t)(wnext - 2) * 32u - (uint32_t)vb; }
    __device__ __forceinline__ bool overrun() const { return (wnext >> 2) > qlast + 2; }
    __device__ __forceinline__ uint32_t window() const { return hi; }
    __device__ __forceinline__ void merge_word() {                 // vb < 32: append nx below the valid bits
        hi |= __funnelshift_rc(nx, 0u, (uint32_t)vb);
        lo = __funnelshift_lc(0u, nx, 32u - (uint32_t)vb);
        vb += 32;
        nx = bswap32(nx2raw); nx2raw = load_raw(wnext); wnext++;
    }
    // The same as `if (vb < 32) merge_word();` as straight-line predicated code: with 32 independent streams per warp
    // some lane merges in almost every step, so a branch would cost every lane the divergent path plus its
    // reconvergence; the freshly loaded word is not touched before the next merge (no wait on the LDS).
    __device__ __forceinline__ void merge_word_predicated() {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b32 t, a;\n\t"
            "setp.lt.s32 p, %4, 32;\n\t"
            "@p shf.r.clamp.b32 t, %2, 0, %4;\n\t"
            "@p or.b32 %0, %0, t;\n\t"
            "@p sub.s32 t, 32, %4;\n\t"
            "@p shf.l.clamp.b32 %1, 0, %2, t;\n\t"
            "@p add.s32 %4, %4, 32;\n\t"
            "@p prmt.b32 %2, %3, 0, 0x0123;\n\t"
            "@p shl.b32 a, %5, 2;\n\t"
            "@p and.b32 a, a, 252;\n\t"
            "@p xor.b32 a, a, %6;\n\t"
            "@p add.u32 a, a, %7;\n\t"
            "@p ld.shared.u32 %3, [a];\n\t"
            "@p add.u32 %5, %5, 1;\n\t"
            "}\n"
            : "+r"(hi), "+r"(lo), "+r"(nx), "+r"(nx2raw), "+r"(vb), "+r"(wnext)
            : "r"(swz), "r"(sbase)
            : "memory");
    }
    __device__ __forceinline__ void consume(uint32_t nb) {         // nb <= 32
        hi = __funnelshift_lc(lo, hi, nb);
        lo = __funnelshift_lc(0u, lo, nb);
        vb -= (int32_t)nb;
        if (vb < 32) merge_word();
    }
    // generic (rare) operations for headers, escapes and over-long codes; callers top up between them
    __device__ __forceinline__ uint32_t get(uint32_t nb) {          // nb in 0..32
        const uint32_t v = __funnelshift_lc(hi, 0u, nb);            // hi >> (32 - nb), 0 for nb == 0
        consume(nb);
        return v;
    }
    __device__ __forceinline__ int32_t get_signed(uint32_t nb) {    // nb in 0..33
        if (nb == 0) return 0;
        if (nb > 32) { consume(nb - 32); nb = 32; }                 // top bits are sign copies for in-range data
        const uint32_t v = get(nb);
        const uint32_t sh = 32 - nb;
        return (int32_t)(v << sh) >> sh;
    }
    __device__ __forceinline__ uint32_t unary() {
        uint32_t q = 0;
        for (;;) {
            if (hi) { const uint32_t z = __clz(hi); consume(z + 1); return q + z; }
            q += 32; consume(32);
            top_up();
            if (overrun()) return q;                                // ran past the frame: corrupt stream
        }
    }
    // zig-zag folded Rice value with parameter k (<= 30), any length (generic path)
    __device__ __forceinline__ uint32_t rice_u(uint32_t k) {
        const uint32_t q = unary();
        const uint32_t low = get(k);
        return (q << k) | low;
    }
    __device__ __forceinline__ void seek(const uint8_t *base, uint64_t bits_forward, uint64_t byte_end) {
        const uint64_t target = bitpos() + bits_forward;
        cp_async_wait<0>();                                          // nothing may land in the ring after re-initialisation
        init(sbase, swz >> 4, base, target, byte_end);
    }
};

__global__ void __launch_bounds__(kDecThreads)
k_skim_subframes(const uint8_t *__restrict__ bytes, const DecStreamDev *__restrict__ streams, uint32_t n_streams,
                 uint32_t channels, uint32_t bps, uint32_t blocksize, uint32_t total_frames,
                 const unsigned long long *__restrict__ frame_pos, uint32_t *__restrict__ sub_bitoff,
                 uint8_t *__restrict__ frame_chassign, uint32_t *__restrict__ status, uint32_t lanes_per_warp) {
    __shared__ __align__(256) uint4 s_ring[kSkimRing * kDecThreads];
    // Only `lanes_per_warp` lanes of every warp take a frame: the walk is a long dependent chain, and there are far
    // fewer frames than the machine has thread slots, so spreading them over more warps buys latency hiding (and
    // less divergence per warp) for issue slots that would otherwise idle.
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t f = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes_per_warp + lane;
    // No lane leaves before the walk is over: the loop below re-converges the whole warp at every trip
    // (__any_sync).  Without that, lanes in different phases (batch / partition change / subframe header) ran
    // their phases one after the other and a warp executed 3.3x the batches a single lane needs
    // (profiles/r01_ncu_skim_v5.txt).
    bool done = !(lane < lanes_per_warp && f < total_frames) || channels <= 1;
    DecStreamDev st;
    FrameLoc L; L.ok = false; L.start = L.end = 0; L.k = 0; L.n = 0; L.stream = 0;
    FrameHdr h; h.header_bytes = 0; h.ch_assign = 0;
    uint32_t *off_out = sub_bitoff + (size_t)(done ? 0 : f) * channels;
    uint64_t frame_bit0 = 0;
    BitReader br;
    br.gq = (const uint4 *)bytes; br.sbase = (uint32_t)__cvta_generic_to_shared(s_ring + threadIdx.x * kSkimRing); br.swz = (lane & 15u) << 4;
    br.wnext = 2; br.cissue = 0; br.qlast = 0; br.hi = br.lo = br.nx = br.nx2raw = 0; br.vb = 64;
    if (lane < lanes_per_warp && f < total_frames) {
        L = locate_frame(streams, n_streams, blocksize, f, frame_pos, &st);
        for (uint32_t c = 0; c < channels; c++) off_out[c] = 0;
        if (!L.ok) { atomicAdd(&status[0], 1u); done = true; }
        else if (!parse_frame_header(bytes + L.start, L.end - L.start, 0, bps, &h)) { atomicAdd(&status[2], 1u); done = true; }
        else {
            frame_chassign[f] = (uint8_t)h.ch_assign;
            frame_bit0 = L.start * 8;
            off_out[0] = h.header_bytes * 8;
            if (!done) br.init(br.sbase, lane, bytes, frame_bit0 + (uint64_t)h.header_bytes * 8, L.end);
        }
    }
    const uint32_t n = L.n;
    bool err = false, in_res = false;
    uint32_t c = 0, left = 0, parts_left = 0, k = 0, plen = 4, esc = 15, psize = 0;
    while (__any_sync(0xFFFFFFFFu, !done)) {
        if (done) continue;
        br.top_up();
        if (left >= (uint32_t)kSkimBatch) {
            // ---- a full batch of Rice codes without data-dependent branches; a code longer than 32 bits (long unary
            // run or corrupt data) is detected once per batch and the batch is then redone one code at a time ----
            const BitReader snap = br;
            const uint32_t k1 = k + 1;
            uint32_t maxlen = 0;
#pragma unroll
            for (int i = 0; i < kSkimBatch; i++) {
                const uint32_t len = (uint32_t)__clz(br.hi) + k1;
                maxlen = max(maxlen, len);
                br.hi = __funnelshift_lc(br.lo, br.hi, len);
                br.lo = __funnelshift_lc(0u, br.lo, len);
                br.vb -= (int32_t)len;
                br.merge_word_predicated();
            }
            if (maxlen <= 32) { left -= kSkimBatch; continue; }
            br = snap;
        }
        if (left) {
            // ---- one batch of Rice codes, predicated on the lane still having codes in this partition ----
            const uint32_t m = left < (uint32_t)kSkimBatch ? left : (uint32_t)kSkimBatch;
            const uint32_t k1 = k + 1;
#pragma unroll
            for (int i = 0; i < kSkimBatch; i++) {
                const uint32_t w = br.window();
                uint32_t len = (uint32_t)__clz(w) + k1;
                const bool active = (uint32_t)i < m;
                if (active && len > 32) {                    // rare: long unary run (or a corrupt stream)
                    (void)br.unary();
                    len = k;
                    if (br.overrun()) { err = true; len = 0; }
                }
                br.consume(active ? len : 0u);
            }
            left -= m;
            if (err) done = true;
            continue;
        }
        if (in_res) {
            // partition exhausted: next partition, or the subframe is finished
            for (;;) {
                if (--parts_left == 0) { in_res = false; break; }
                left = psize;
                k = br.get(plen);
                if (k == esc) { const uint32_t raw = br.get(5); br.seek(bytes, (uint64_t)raw * left, L.end); left = 0; }
                if (left) break;
            }
            if (!in_res) { c++; off_out[c] = (uint32_t)(br.bitpos() - frame_bit0); }
            continue;
        }
        if (err || c + 1 >= channels) { done = true; continue; }
        {
            uint32_t sbps = bps;
            if ((h.ch_assign == 8 && c == 1) || (h.ch_assign == 9 && c == 0) || (h.ch_assign == 10 && c == 1)) sbps++;
            const uint32_t hd = br.get(8);
            const uint32_t t = (hd >> 1) & 0x3F;
            uint32_t wasted = 0;
            if (hd & 1) wasted = br.unary() + 1;
            if ((hd & 0x80) || wasted >= sbps) { err = true; done = true; continue; }
            sbps -= wasted;
            uint32_t order = 0;
            bool coded = false;
            if (t == 0) br.seek(bytes, sbps, L.end);
            else if (t == 1) br.seek(bytes, (uint64_t)sbps * n, L.end);
            else if (t >= 8 && t <= 12) { order = t - 8; coded = true; }
            else if (t >= 32) { order = t - 31; coded = true; }
            else { err = true; done = true; continue; }
            if (coded) {
                if (order > n) { err = true; done = true; continue; }
                br.seek(bytes, (uint64_t)order * sbps, L.end);
                if (t >= 32) {
                    const uint32_t prec = br.get(4) + 1;
                    if (prec == 16) { err = true; done = true; continue; }
                    br.seek(bytes, 5 + (uint64_t)order * prec, L.end);
                }
                const uint32_t m = br.get(2);
                const uint32_t po = br.get(4);
                plen = m ? 5u : 4u; esc = m ? 31u : 15u;
                psize = n >> po;
                if (m > 1 || (po > 0 && (n & ((1u << po) - 1))) || psize < order) { err = true; done = true; continue; }
                parts_left = 1u << po;
                left = psize - order;
                in_res = true;
                // settle on the first partition that actually holds Rice-coded residuals
                for (;;) {
                    k = br.get(plen);
                    if (k == esc) { const uint32_t raw = br.get(5); br.seek(bytes, (uint64_t)raw * left, L.end); left = 0; }
                    if (left) break;
                    if (--parts_left == 0) { in_res = false; break; }
                    left = psize;
                    if (psize == 0) { err = true; break; }
                }
            }
            if (!in_res) { c++; off_out[c] = (uint32_t)(br.bitpos() - frame_bit0); }
        }
        if (err) done = true;
    }
    const bool mine = lane < lanes_per_warp && f < total_frames && channels > 1 && L.ok && h.header_bytes != 0;
    if (mine && !err && br.bitpos() > L.end * 8) err = true;
    if (mine && err) {
        atomicAdd(&status[2], 1u);
        for (uint32_t q = 0; q < channels; q++) off_out[q] = 0;
    }
}

// ---- subframe decode -------------------------------------------------------------------------------
template <int MAXORD, bool WIDE>
__device__ __forceinline__ int32_t lpc_predict(const int32_t (&cf)[MAXORD], const int32_t *Hj, int shift) {
    // Hj points at the newest history sample; taps walk backwards (static indices after unrolling)
    if (WIDE) {
        long long acc = 0;
#pragma unroll
        for (int q = 0; q < MAXORD; q++) acc += (long long)cf[q] * (long long)Hj[-q];
        return (int32_t)(acc >> shift);
    } else {
        int32_t acc = 0;
#pragma unroll
        for (int q = 0; q < MAXORD; q++) acc += cf[q] * Hj[-q];
        return acc >> shift;
    }
}

struct SubCtx {
    BitReader br;
    int32_t *dst;            // first sample of this subframe in the planar audio buffer
    uint32_t n, order, sbps, wasted, type;
    bool aligned16;
    bool err;
};

constexpr int kDecBatch = 8;      // samples decoded per branch-free batch

// FIXED / LPC subframe body.  Called by ALL lanes of the warp (`active` false for lanes without a predictive
// subframe): the sample loop re-converges the warp at every trip, and a trip is either one batch of kDecBatch
// Rice codes parsed without data-dependent branches (predicated word merges, see BitReader) followed by the
// predictor recursion from the register history, or a single sample through the generic path (escape-coded
// partitions, partition tails, codes longer than 32 bits).
template <int MAXORD, bool WIDE>
__device__ __forceinline__ void decode_predictive(SubCtx &S, bool active) {
    BitReader &br = S.br;
    const uint32_t n = S.n, order = S.order, wasted = S.wasted;
    int32_t *dst = S.dst;
    int32_t H[MAXORD + kDecBatch];
    int32_t cf[MAXORD];
#pragma unroll
    for (int q = 0; q < MAXORD + kDecBatch; q++) H[q] = 0;
#pragma unroll
    for (int q = 0; q < MAXORD; q++) cf[q] = 0;
    int shift = 0;
    uint32_t k = 0, plen = 4, esc = 15, psize = 0, part_left = 0, raw_bits = 0, i = n;
    bool escape = false;
    if (active) {
        // warm-up samples
        for (uint32_t w = 0; w < order; w++) {
            const int32_t v = br.get_signed(S.sbps);
#pragma unroll
            for (int q = 0; q < MAXORD - 1; q++) H[q] = H[q + 1];
            H[MAXORD - 1] = v;
            dst[w] = (int32_t)((uint32_t)v << wasted);
        }
        br.top_up();
        if (S.type == 3) {
            const uint32_t prec = br.get(4) + 1;
            const uint32_t sh = br.get(5);
            if (prec == 16 || (sh & 16)) S.err = true;
            shift = (int)sh;
#pragma unroll
            for (int q = 0; q < MAXORD; q++) if ((uint32_t)q < order) cf[q] = br.get_signed(prec);
            br.top_up();
        } else {
            if (order >= 1) cf[0] = order == 1 ? 1 : order == 2 ? 2 : order == 3 ? 3 : 4;
            if (MAXORD > 1 && order >= 2) cf[1] = order == 2 ? -1 : order == 3 ? -3 : -6;
            if (MAXORD > 2 && order >= 3) cf[2] = order == 3 ? 1 : 4;
            if (MAXORD > 3 && order >= 4) cf[3] = -1;
        }
        const uint32_t m = br.get(2);
        const uint32_t po = br.get(4);
        plen = m ? 5u : 4u; esc = m ? 31u : 15u;
        psize = n >> po;
        if (m > 1 || (po > 0 && (n & ((1u << po) - 1))) || psize < order) S.err = true;
        if (!S.err) {
            k = br.get(plen);
            escape = (k == esc);
            raw_bits = escape ? br.get(5) : 0;
            part_left = psize - order;
            i = order;
            while (part_left == 0 && i < n) {
                if (psize == 0) { S.err = true; break; }
                k = br.get(plen); escape = (k == esc); raw_bits = escape ? br.get(5) : 0; part_left = psize;
            }
            if (S.err) i = n;
        }
    }
    while (__any_sync(0xFFFFFFFFu, i < n)) {
        if (i >= n) continue;
        br.top_up();
        bool did = false;
        if (!escape && part_left >= (uint32_t)kDecBatch && (i & 3u) == 0) {       // (i & 3): keep the 16-byte stores aligned
            // ---- batch: kDecBatch codes, no data-dependent branch; redone sample by sample if one is longer than 32 bits ----
            const BitReader snap = br;
            const uint32_t k1 = k + 1;
            uint32_t maxlen = 0, u[kDecBatch];
#pragma unroll
            for (int j = 0; j < kDecBatch; j++) {
                const uint32_t z = (uint32_t)__clz(br.hi);
                const uint32_t len = z + k1;
                maxlen = max(maxlen, len);
                const uint32_t t = __funnelshift_lc(0u, br.hi, z + 1);      // hi << (z+1), 0 when z+1 == 32
                u[j] = (z << k) | __funnelshift_lc(t, 0u, k);               // | t >> (32-k), 0 when k == 0
                br.hi = __funnelshift_lc(br.lo, br.hi, len);
                br.lo = __funnelshift_lc(0u, br.lo, len);
                br.vb -= (int32_t)len;
                br.merge_word_predicated();
            }
            if (maxlen <= 32) {
#pragma unroll
                for (int j = 0; j < kDecBatch; j++)
                    H[MAXORD + j] = unzigzag(u[j]) + lpc_predict<MAXORD, WIDE>(cf, &H[MAXORD - 1 + j], shift);
                if (S.aligned16) {
#pragma unroll
                    for (int q = 0; q < kDecBatch / 4; q++)
                        *reinterpret_cast<int4 *>(dst + i + 4 * q) =
                            make_int4((int32_t)((uint32_t)H[MAXORD + 4 * q] << wasted), (int32_t)((uint32_t)H[MAXORD + 4 * q + 1] << wasted),
                                      (int32_t)((uint32_t)H[MAXORD + 4 * q + 2] << wasted), (int32_t)((uint32_t)H[MAXORD + 4 * q + 3] << wasted));
                } else {
#pragma unroll
                    for (int j = 0; j < kDecBatch; j++) dst[i + j] = (int32_t)((uint32_t)H[MAXORD + j] << wasted);
                }
#pragma unroll
                for (int q = 0; q < MAXORD; q++) H[q] = H[q + kDecBatch];
                part_left -= kDecBatch; i += kDecBatch;
                did = true;
            } else {
                br = snap;
            }
        }
        if (!did) {
            const int32_t r = escape ? br.get_signed(raw_bits) : unzigzag(br.rice_u(k));
            const int32_t v = r + lpc_predict<MAXORD, WIDE>(cf, &H[MAXORD - 1], shift);
#pragma unroll
            for (int q = 0; q < MAXORD - 1; q++) H[q] = H[q + 1];
            H[MAXORD - 1] = v;
            dst[i] = (int32_t)((uint32_t)v << wasted);
            part_left--; i++;
            if (br.overrun()) { S.err = true; i = n; }
        }
        if (part_left == 0 && i < n) {
            k = br.get(plen); escape = (k == esc); raw_bits = escape ? br.get(5) : 0; part_left = psize;
        }
    }
}

#ifndef FRB_DEC_MINB
#define FRB_DEC_MINB 4
#endif
template <bool BIGORDER>
__global__ void __launch_bounds__(kDecThreads, BIGORDER ? 1 : FRB_DEC_MINB)
k_decode_subframes(const uint8_t *__restrict__ bytes, const DecStreamDev *__restrict__ streams, uint32_t n_streams,
                   uint32_t channels, uint32_t bps, uint32_t blocksize, uint32_t total_frames,
                   const unsigned long long *__restrict__ frame_pos, const uint32_t *__restrict__ sub_bitoff,
                   int32_t *__restrict__ audio, uint8_t *__restrict__ frame_chassign, uint32_t *__restrict__ status) {
    __shared__ __align__(256) uint4 s_ring[kSkimRing * kDecThreads];
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t total_sub = total_frames * channels;
    bool alive = s < total_sub;
    const uint32_t f = alive ? s / channels : 0, c = alive ? s - f * channels : 0;
    SubCtx S;
    S.err = false; S.type = 0; S.order = 0; S.n = 0; S.wasted = 0; S.sbps = bps; S.dst = audio; S.aligned16 = false;
    S.br.gq = (const uint4 *)bytes; S.br.sbase = (uint32_t)__cvta_generic_to_shared(s_ring + threadIdx.x * kSkimRing);
    S.br.swz = (lane & 15u) << 4; S.br.wnext = 2; S.br.cissue = 0; S.br.qlast = 0; S.br.hi = S.br.lo = S.br.nx = S.br.nx2raw = 0; S.br.vb = 64;
    FrameLoc L; L.ok = false; L.start = L.end = 0; L.k = 0; L.n = 0;
    uint32_t hdr_bytes = 0;
    if (alive) {
        DecStreamDev st;
        L = locate_frame(streams, n_streams, blocksize, f, frame_pos, &st);
        if (!L.ok) { alive = false; if (c == 0 && !sub_bitoff) atomicAdd(&status[0], 1u); }
        else {
            uint32_t ch_assign = 0;
            uint64_t bit0;
            if (sub_bitoff) {
                const uint32_t off = sub_bitoff[s];
                if (off == 0) alive = false;               // the skim pass already counted the error
                bit0 = L.start * 8 + off;
                ch_assign = frame_chassign[f];
            } else {
                FrameHdr h;
                if (!parse_frame_header(bytes + L.start, L.end - L.start, 0, bps, &h)) { alive = false; atomicAdd(&status[2], 1u); }
                else { hdr_bytes = h.header_bytes; ch_assign = h.ch_assign; frame_chassign[f] = (uint8_t)h.ch_assign; }
                bit0 = (L.start + hdr_bytes) * 8;
            }
            if (alive) {
                S.br.init(S.br.sbase, lane, bytes, bit0, L.end);
                S.n = L.n;
                const int64_t idx = st.audio_base + (int64_t)c * (int64_t)st.n_samples + (int64_t)L.k * blocksize;
                S.dst = audio + idx;
                S.aligned16 = ((reinterpret_cast<uintptr_t>(S.dst) & 15u) == 0);
                if ((ch_assign == 8 && c == 1) || (ch_assign == 9 && c == 0) || (ch_assign == 10 && c == 1)) S.sbps++;
                const uint32_t hd = S.br.get(8);
                if (hd & 0x80) S.err = true;
                const uint32_t t = (hd >> 1) & 0x3F;
                if (hd & 1) S.wasted = S.br.unary() + 1;
                if (S.wasted >= S.sbps) { S.err = true; S.wasted = 0; }
                S.sbps -= S.wasted;
                if (t == 0) S.type = 0;
                else if (t == 1) S.type = 1;
                else if (t >= 8 && t <= 12) { S.type = 2; S.order = t - 8; }
                else if (t >= 32) { S.type = 3; S.order = t - 31; }
                else S.err = true;
                if (S.order > S.n) S.err = true;
            }
        }
    }
    const bool run = alive && !S.err;
    // warp-uniform code path: taps padded to the largest order in the warp, 64-bit MACs if any lane needs them
    uint32_t my_ord = (run && S.type >= 2) ? S.order : 0;
    if (!BIGORDER && my_ord > 12) { my_ord = 0; }
    const uint32_t cls = __reduce_max_sync(0xFFFFFFFFu, my_ord);
    bool wide_lane = false;
    if (run && S.type >= 2) {
        // libFLAC's rule: 32-bit arithmetic is exact when bps + precision + ilog2(order) <= 32 (fixed: bps + order)
        wide_lane = S.sbps + (S.type == 3 ? 15u + (uint32_t)(31 - __clz(S.order | 1u)) : S.order) > 32u;
    }
    const bool wide = __any_sync(0xFFFFFFFFu, wide_lane);
    if (run) {
        if (S.type == 0) {
            const int32_t v = (int32_t)((uint32_t)S.br.get_signed(S.sbps) << S.wasted);
            for (uint32_t i = 0; i < S.n; i++) S.dst[i] = v;
        } else if (S.type == 1) {
            for (uint32_t i = 0; i < S.n; i++) {
                if ((i & 3u) == 0) S.br.top_up();
                S.dst[i] = (int32_t)((uint32_t)S.br.get_signed(S.sbps) << S.wasted);
            }
        } else if (!BIGORDER && S.order > 12) {
            S.err = true;
            atomicAdd(&status[4], 1u);
        }
    }
    {
        // every lane of the warp enters the same instantiation (the sample loop is warp-synchronous)
        const bool act = run && S.type >= 2 && (BIGORDER || S.order <= 12);
        if (BIGORDER) decode_predictive<32, true>(S, act);
        else if (cls <= 4) { if (wide) decode_predictive<4, true>(S, act); else decode_predictive<4, false>(S, act); }
        else if (cls <= 8) { if (wide) decode_predictive<8, true>(S, act); else decode_predictive<8, false>(S, act); }
        else { if (wide) decode_predictive<12, true>(S, act); else decode_predictive<12, false>(S, act); }
    }
    if (alive) {
        if (!S.err && c + 1 == channels) {
            // the last subframe must end (after byte padding) exactly 2 bytes before the next frame
            const uint64_t bits = S.br.bitpos() - L.start * 8;
            if (L.start + ((bits + 7) >> 3) + 2 != L.end) S.err = true;
        }
        if (S.err) atomicAdd(&status[2], 1u);
        else if (c + 1 == channels) atomicAdd(&status[3], 1u);
    }
}

#include "frb_crc16.cuh"   // warp-cooperative CRC-16 + k_crc16_frames
