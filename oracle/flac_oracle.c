/*
 * flac_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * A plain-C, single-threaded restatement of the codec arithmetic that the
 * reference (yharby/flac-raster) obtains from its third-party dependency
 * pyflac==3.0.0 (uv.lock:844-852), which bundles a prebuilt libFLAC 1.4.3
 * (docs/sonos-pyflac.txt:150, :2431).  libFLAC's C source is NOT present in
 * /root/reference, so this file restates the *published* algorithm:
 *   - bitstream/decoder: RFC 9639 and the constants in
 *     docs/sonos-pyflac.txt:3394-4431 (format.h text);
 *   - encoder decision procedure: libFLAC 1.4.3 process_subframe_ /
 *     find_best_partition_order_ / set_partitioned_rice_ /
 *     FLAC__lpc_* / FLAC__fixed_compute_best_predictor / FLAC__window_tukey,
 *     with the preset table of docs/sonos-pyflac.txt:6910-6935.
 * It is anchored on the reference's own call sites
 *   converter.py:139-154 (StreamEncoder, blocksize=4096, level 5)
 *   converter.py:181-182 (FileDecoder.process)
 * and pinned on the reference's golden vectors test_data/sample_rgb.flac
 * (+ sample_rgb.tif) and test_data/sample_dem.flac (see tests/test_oracle_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product path never does.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o _build/libflac_oracle.so flac_oracle.c -lm
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#ifndef M_LN2
#define M_LN2 0.69314718055994530942
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define FO_MAX_CHANNELS 8
#define FO_MAX_LPC_ORDER 32
#define FO_MAX_BLOCK 65535
#define FO_DESC_PARAMS 64

/* ------------------------------------------------------------------ */
/* public structs (mirrored with ctypes in oracle/flac_oracle.py)      */
/* ------------------------------------------------------------------ */
typedef struct {
    uint32_t sample_rate, channels, bps;
    uint32_t min_blocksize, max_blocksize;
    uint32_t min_framesize, max_framesize;
    uint64_t total_samples_streaminfo;
    uint64_t samples_decoded;      /* per channel */
    uint32_t n_frames;
    uint32_t first_frame_offset;   /* byte offset of the first audio frame */
    uint32_t crc8_errors, crc16_errors;
    uint64_t bytes_consumed;       /* where decoding stopped (next "fLaC" or EOF) */
    uint8_t  md5[16];
    uint32_t vorbis_offset, vorbis_length; /* VORBIS_COMMENT block body */
} fo_stream_info;

typedef struct {
    uint32_t frame, channel;
    uint32_t type;        /* 0 CONSTANT 1 VERBATIM 2 FIXED 3 LPC */
    uint32_t order, wasted, precision;
    int32_t  shift;
    int32_t  coefs[FO_MAX_LPC_ORDER];
    uint32_t method, partition_order;
    uint32_t params[FO_DESC_PARAMS];   /* first 64 rice parameters (0xFF.. = escape) */
    uint32_t n_escape;
    uint32_t nbits;       /* total bits of the subframe */
    uint32_t blocksize;
    uint32_t ch_assign;
    uint32_t frame_bytes;
    uint32_t frame_offset;
} fo_subframe_desc;

/* ------------------------------------------------------------------ */
/* CRC                                                                 */
/* ------------------------------------------------------------------ */
static uint8_t  crc8_tab[256];
static uint16_t crc16_tab[256];
static int crc_ready = 0;
static void crc_init(void) {
    if (crc_ready) return;
    for (int i = 0; i < 256; i++) {
        uint8_t c = (uint8_t)i;
        for (int b = 0; b < 8; b++) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
        crc8_tab[i] = c;
        uint16_t d = (uint16_t)(i << 8);
        for (int b = 0; b < 8; b++) d = (d & 0x8000) ? (uint16_t)((d << 1) ^ 0x8005) : (uint16_t)(d << 1);
        crc16_tab[i] = d;
    }
    crc_ready = 1;
}
uint8_t fo_crc8(const uint8_t *p, size_t n) {
    crc_init();
    uint8_t c = 0;
    while (n--) c = crc8_tab[c ^ *p++];
    return c;
}
uint16_t fo_crc16(const uint8_t *p, size_t n) {
    crc_init();
    uint16_t c = 0;
    while (n--) c = (uint16_t)((c << 8) ^ crc16_tab[(c >> 8) ^ *p++]);
    return c;
}

/* ------------------------------------------------------------------ */
/* bit reader (MSB first)                                              */
/* ------------------------------------------------------------------ */
typedef struct { const uint8_t *p; size_t len; size_t bitpos; int err; } br_t;
static inline uint32_t br_bit(br_t *b) {
    size_t byte = b->bitpos >> 3;
    if (byte >= b->len) { b->err = 1; return 0; }
    uint32_t v = (b->p[byte] >> (7 - (b->bitpos & 7))) & 1u;
    b->bitpos++;
    return v;
}
static inline uint64_t br_bits(br_t *b, int n) {   /* n <= 57 */
    uint64_t v = 0;
    while (n > 0) {
        size_t byte = b->bitpos >> 3;
        if (byte >= b->len) { b->err = 1; return 0; }
        int avail = 8 - (int)(b->bitpos & 7);
        int take = n < avail ? n : avail;
        uint32_t cur = b->p[byte] & ((1u << avail) - 1u);
        v = (v << take) | (cur >> (avail - take));
        b->bitpos += (size_t)take;
        n -= take;
    }
    return v;
}
static inline int64_t br_sbits(br_t *b, int n) {
    if (n == 0) return 0;
    uint64_t v = br_bits(b, n);
    uint64_t m = 1ull << (n - 1);
    return (int64_t)((v ^ m) - m);
}
static inline uint32_t br_unary(br_t *b) {       /* count zeros before a 1 */
    uint32_t q = 0;
    while (!b->err && br_bit(b) == 0) q++;
    return q;
}

/* ------------------------------------------------------------------ */
/* decoder                                                             */
/* ------------------------------------------------------------------ */
static const uint32_t rate_tab[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
static const uint32_t bps_tab[8]   = {0, 8, 12, 0, 16, 20, 24, 32};

typedef struct {
    uint32_t blocksize, sample_rate, channels, ch_assign, bps;
    uint64_t number;
    int variable;
    uint32_t header_bytes;
} frame_hdr;

/* returns 0 ok, <0 on invalid header */
static int parse_frame_header(const uint8_t *p, size_t len, const fo_stream_info *si, frame_hdr *h) {
    if (len < 5) return -1;
    if (p[0] != 0xFF || (p[1] & 0xFE) != 0xF8) return -1;
    h->variable = p[1] & 1;
    uint32_t bsc = p[2] >> 4, src = p[2] & 15, chc = p[3] >> 4, bpc = (p[3] >> 1) & 7;
    if (p[3] & 1) return -1;
    if (bsc == 0 || src == 15 || bpc == 3 || chc > 10) return -1;
    size_t i = 4;
    /* UTF-8 style number */
    uint32_t b0 = p[i++];
    uint64_t num; int extra;
    if (b0 < 0x80) { num = b0; extra = 0; }
    else if ((b0 & 0xE0) == 0xC0) { num = b0 & 0x1F; extra = 1; }
    else if ((b0 & 0xF0) == 0xE0) { num = b0 & 0x0F; extra = 2; }
    else if ((b0 & 0xF8) == 0xF0) { num = b0 & 0x07; extra = 3; }
    else if ((b0 & 0xFC) == 0xF8) { num = b0 & 0x03; extra = 4; }
    else if ((b0 & 0xFE) == 0xFC) { num = b0 & 0x01; extra = 5; }
    else if (b0 == 0xFE) { num = 0; extra = 6; }
    else return -1;
    if (!h->variable && extra > 5) return -1;
    for (int k = 0; k < extra; k++) {
        if (i >= len) return -1;
        if ((p[i] & 0xC0) != 0x80) return -1;
        num = (num << 6) | (p[i++] & 0x3F);
    }
    h->number = num;
    if (bsc == 1) h->blocksize = 192;
    else if (bsc <= 5) h->blocksize = 576u << (bsc - 2);
    else if (bsc == 6) { if (i >= len) return -1; h->blocksize = (uint32_t)p[i++] + 1; }
    else if (bsc == 7) { if (i + 1 >= len) return -1; h->blocksize = (((uint32_t)p[i] << 8) | p[i + 1]) + 1; i += 2; }
    else h->blocksize = 256u << (bsc - 8);
    if (src == 0) h->sample_rate = si ? si->sample_rate : 0;
    else if (src <= 11) h->sample_rate = rate_tab[src];
    else if (src == 12) { if (i >= len) return -1; h->sample_rate = (uint32_t)p[i++] * 1000u; }
    else if (src == 13) { if (i + 1 >= len) return -1; h->sample_rate = ((uint32_t)p[i] << 8) | p[i + 1]; i += 2; }
    else { if (i + 1 >= len) return -1; h->sample_rate = (((uint32_t)p[i] << 8) | p[i + 1]) * 10u; i += 2; }
    h->ch_assign = chc;
    h->channels = chc < 8 ? chc + 1 : 2;
    h->bps = bpc == 0 ? (si ? si->bps : 0) : bps_tab[bpc];
    if (i >= len) return -1;
    if (fo_crc8(p, i) != p[i]) return -2;
    h->header_bytes = (uint32_t)(i + 1);
    return 0;
}

static int decode_residual(br_t *b, int64_t *res, uint32_t blocksize, uint32_t order, fo_subframe_desc *d) {
    uint32_t method = (uint32_t)br_bits(b, 2);
    if (method > 1) return -10;
    uint32_t po = (uint32_t)br_bits(b, 4);
    uint32_t parts = 1u << po;
    int plen = method ? 5 : 4;
    uint32_t esc = method ? 31 : 15;
    if (d) { d->method = method; d->partition_order = po; d->n_escape = 0; }
    if ((blocksize >> po) < order && po > 0) return -11;
    if (po > 0 && (blocksize & (parts - 1))) return -12;
    uint32_t i = order;
    for (uint32_t p = 0; p < parts; p++) {
        uint32_t n = blocksize >> po;
        if (p == 0) { if (n < order) return -13; n -= order; }
        uint32_t k = (uint32_t)br_bits(b, plen);
        if (d && p < FO_DESC_PARAMS) d->params[p] = k;
        if (k == esc) {
            uint32_t raw = (uint32_t)br_bits(b, 5);
            if (d) { d->n_escape++; if (p < FO_DESC_PARAMS) d->params[p] = 0xFF00u | raw; }
            for (uint32_t j = 0; j < n; j++) res[i++] = br_sbits(b, (int)raw);
        } else {
            for (uint32_t j = 0; j < n; j++) {
                uint32_t q = br_unary(b);
                uint64_t u = ((uint64_t)q << k) | (k ? br_bits(b, (int)k) : 0);
                res[i++] = (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
                if (b->err) return -14;
            }
        }
    }
    return b->err ? -14 : 0;
}

static int decode_subframe(br_t *b, int64_t *out, uint32_t blocksize, uint32_t bps, fo_subframe_desc *d) {
    size_t start = b->bitpos;
    if (br_bit(b)) return -20;
    uint32_t type = (uint32_t)br_bits(b, 6);
    uint32_t wasted = 0;
    if (br_bit(b)) wasted = br_unary(b) + 1;
    if (wasted >= bps) return -21;
    bps -= wasted;
    if (d) { memset(d->coefs, 0, sizeof d->coefs); memset(d->params, 0, sizeof d->params);
             d->wasted = wasted; d->order = 0; d->precision = 0; d->shift = 0; d->method = 0; d->partition_order = 0; d->n_escape = 0; }
    if (type == 0) {
        int64_t v = br_sbits(b, (int)bps);
        for (uint32_t i = 0; i < blocksize; i++) out[i] = v;
        if (d) d->type = 0;
    } else if (type == 1) {
        for (uint32_t i = 0; i < blocksize; i++) out[i] = br_sbits(b, (int)bps);
        if (d) d->type = 1;
    } else if (type >= 8 && type <= 12) {
        uint32_t order = type - 8;
        if (order > blocksize) return -22;
        for (uint32_t i = 0; i < order; i++) out[i] = br_sbits(b, (int)bps);
        if (d) { d->type = 2; d->order = order; }
        int rc = decode_residual(b, out, blocksize, order, d);
        if (rc) return rc;
        for (uint32_t i = order; i < blocksize; i++) {
            int64_t p = 0;
            switch (order) {
                case 0: p = 0; break;
                case 1: p = out[i - 1]; break;
                case 2: p = 2 * out[i - 1] - out[i - 2]; break;
                case 3: p = 3 * out[i - 1] - 3 * out[i - 2] + out[i - 3]; break;
                case 4: p = 4 * out[i - 1] - 6 * out[i - 2] + 4 * out[i - 3] - out[i - 4]; break;
            }
            out[i] += p;
        }
    } else if (type >= 32) {
        uint32_t order = type - 31;
        if (order > blocksize) return -23;
        for (uint32_t i = 0; i < order; i++) out[i] = br_sbits(b, (int)bps);
        uint32_t prec = (uint32_t)br_bits(b, 4);
        if (prec == 15) return -24;
        prec += 1;
        int32_t shift = (int32_t)br_sbits(b, 5);
        if (shift < 0) return -25;
        int64_t c[FO_MAX_LPC_ORDER];
        for (uint32_t j = 0; j < order; j++) c[j] = br_sbits(b, (int)prec);
        if (d) { d->type = 3; d->order = order; d->precision = prec; d->shift = shift;
                 for (uint32_t j = 0; j < order; j++) d->coefs[j] = (int32_t)c[j]; }
        int rc = decode_residual(b, out, blocksize, order, d);
        if (rc) return rc;
        for (uint32_t i = order; i < blocksize; i++) {
            int64_t s = 0;
            for (uint32_t j = 0; j < order; j++) s += c[j] * out[i - 1 - j];
            out[i] += (s >> shift);
        }
    } else return -26;
    if (wasted) for (uint32_t i = 0; i < blocksize; i++) out[i] = (int64_t)((uint64_t)out[i] << wasted);
    if (d) d->nbits = (uint32_t)(b->bitpos - start);
    return b->err ? -27 : 0;
}

/* Parse the metadata chain. Returns offset of the first frame or <0. */
static long parse_metadata(const uint8_t *data, size_t len, fo_stream_info *si) {
    if (len < 8 || memcmp(data, "fLaC", 4)) return -1;
    size_t pos = 4; int last = 0, have_si = 0;
    si->vorbis_offset = si->vorbis_length = 0;
    while (!last) {
        if (pos + 4 > len) return -2;
        last = data[pos] >> 7;
        uint32_t type = data[pos] & 0x7F;
        uint32_t blen = ((uint32_t)data[pos + 1] << 16) | ((uint32_t)data[pos + 2] << 8) | data[pos + 3];
        pos += 4;
        if (pos + blen > len) return -3;
        if (type == 0) {
            if (blen < 34) return -4;
            const uint8_t *s = data + pos;
            si->min_blocksize = ((uint32_t)s[0] << 8) | s[1];
            si->max_blocksize = ((uint32_t)s[2] << 8) | s[3];
            si->min_framesize = ((uint32_t)s[4] << 16) | ((uint32_t)s[5] << 8) | s[6];
            si->max_framesize = ((uint32_t)s[7] << 16) | ((uint32_t)s[8] << 8) | s[9];
            si->sample_rate = ((uint32_t)s[10] << 12) | ((uint32_t)s[11] << 4) | (s[12] >> 4);
            si->channels = ((s[12] >> 1) & 7) + 1;
            si->bps = (((uint32_t)(s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            si->total_samples_streaminfo = ((uint64_t)(s[13] & 15) << 32) | ((uint64_t)s[14] << 24) | ((uint64_t)s[15] << 16) | ((uint64_t)s[16] << 8) | s[17];
            memcpy(si->md5, s + 18, 16);
            have_si = 1;
        } else if (type == 4) {
            si->vorbis_offset = (uint32_t)pos; si->vorbis_length = blen;
        } else if (type == 127) return -5;
        pos += blen;
    }
    if (!have_si) return -6;
    return (long)pos;
}

/*
 * Decode one FLAC stream held in memory. `out` receives interleaved int32
 * samples (sample-major, channel-minor: the layout pyflac hands back,
 * docs/sonos-pyflac.txt:1809-1854). Frames are found sequentially; decoding
 * stops at EOF or when another "fLaC" marker follows a frame (legacy
 * concatenated --spatial files, spatial_encoder.py:238-241).
 * Returns 0 on success, negative on a hard error.
 */
int fo_decode_stream(const uint8_t *data, size_t len, int32_t *out, size_t out_cap_samples,
                     fo_stream_info *si, fo_subframe_desc *descs, size_t desc_cap, size_t *n_descs) {
    memset(si, 0, sizeof *si);
    long pos0 = parse_metadata(data, len, si);
    if (pos0 < 0) return (int)pos0;
    si->first_frame_offset = (uint32_t)pos0;
    size_t pos = (size_t)pos0, nd = 0;
    uint64_t done = 0;
    int64_t (*chbuf)[FO_MAX_BLOCK + 1] = malloc(sizeof(int64_t) * FO_MAX_CHANNELS * (FO_MAX_BLOCK + 1));
    int ret = 0;
    if (!chbuf) return -99;
    while (pos < len) {
        if (len - pos >= 4 && !memcmp(data + pos, "fLaC", 4)) break;
        frame_hdr h;
        int rc = parse_frame_header(data + pos, len - pos, si, &h);
        if (rc == -2) { si->crc8_errors++; { ret = -30; goto done; } }
        if (rc) { ret = -31; goto done; }
        if (h.channels != si->channels || h.bps != si->bps) { ret = -32; goto done; }
        br_t b = { data + pos, len - pos, (size_t)h.header_bytes * 8, 0 };
        for (uint32_t c = 0; c < h.channels; c++) {
            uint32_t bps = h.bps;
            if ((h.ch_assign == 8 && c == 1) || (h.ch_assign == 9 && c == 0) || (h.ch_assign == 10 && c == 1)) bps++;
            fo_subframe_desc *d = (descs && nd < desc_cap) ? &descs[nd] : NULL;
            rc = decode_subframe(&b, chbuf[c], h.blocksize, bps, d);
            if (rc) { ret = rc; goto done; }
            if (d) { d->frame = si->n_frames; d->channel = c; d->blocksize = h.blocksize; d->ch_assign = h.ch_assign; d->frame_offset = (uint32_t)pos; nd++; }
        }
        b.bitpos = (b.bitpos + 7) & ~(size_t)7;
        size_t fbytes = b.bitpos >> 3;
        if (fbytes + 2 > len - pos) { ret = -33; goto done; }
        uint16_t crc = (uint16_t)(((uint16_t)data[pos + fbytes] << 8) | data[pos + fbytes + 1]);
        if (fo_crc16(data + pos, fbytes) != crc) { si->crc16_errors++; { ret = -34; goto done; } }
        fbytes += 2;
        for (size_t k = nd; k-- > 0 && descs && k < desc_cap && descs[k].frame == si->n_frames; ) descs[k].frame_bytes = (uint32_t)fbytes;
        /* undo inter-channel decorrelation */
        if (h.ch_assign == 8) for (uint32_t i = 0; i < h.blocksize; i++) chbuf[1][i] = chbuf[0][i] - chbuf[1][i];
        else if (h.ch_assign == 9) for (uint32_t i = 0; i < h.blocksize; i++) chbuf[0][i] = chbuf[0][i] + chbuf[1][i];
        else if (h.ch_assign == 10) for (uint32_t i = 0; i < h.blocksize; i++) {
            int64_t m = chbuf[0][i], s = chbuf[1][i];
            m = (int64_t)(((uint64_t)m << 1) | (uint64_t)(s & 1));
            chbuf[0][i] = (m + s) >> 1; chbuf[1][i] = (m - s) >> 1;
        }
        if (out) {
            if ((done + h.blocksize) * h.channels > out_cap_samples) { ret = -35; goto done; }
            for (uint32_t i = 0; i < h.blocksize; i++)
                for (uint32_t c = 0; c < h.channels; c++)
                    out[(done + i) * h.channels + c] = (int32_t)chbuf[c][i];
        }
        done += h.blocksize;
        si->n_frames++;
        pos += fbytes;
    }
done:
    si->samples_decoded = done;
    si->bytes_consumed = pos;
    if (n_descs) *n_descs = nd;
    free(chbuf);
    return ret;
}

/* ------------------------------------------------------------------ */
/* bit writer                                                          */
/* ------------------------------------------------------------------ */
typedef struct { uint8_t *p; size_t cap; size_t bitpos; int err; } bw_t;
static inline void bw_bits(bw_t *w, uint64_t v, int n) {   /* n <= 64 */
    while (n > 0) {
        size_t byte = w->bitpos >> 3;
        if (byte >= w->cap) { w->err = 1; return; }
        int room = 8 - (int)(w->bitpos & 7);
        int take = n < room ? n : room;
        uint32_t chunk = (uint32_t)((v >> (n - take)) & ((1u << take) - 1u));
        if (room == 8) w->p[byte] = 0;
        w->p[byte] |= (uint8_t)(chunk << (room - take));
        w->bitpos += (size_t)take;
        n -= take;
    }
}
static inline void bw_unary(bw_t *w, uint32_t q) {   /* q zeros then a one */
    while (q >= 32) { bw_bits(w, 0, 32); q -= 32; }
    bw_bits(w, 1, (int)q + 1);
}
static inline void bw_rice(bw_t *w, int32_t v, uint32_t k) {
    uint32_t u = ((uint32_t)v << 1) ^ (uint32_t)(v >> 31);
    bw_unary(w, u >> k);
    if (k) bw_bits(w, u & ((1u << k) - 1u), (int)k);
}

/* ------------------------------------------------------------------ */
/* encoder: libFLAC 1.4.3 decision procedure, restated                 */
/* ------------------------------------------------------------------ */
typedef struct {
    int do_mid_side, loose_mid_side;
    int max_lpc_order;
    int min_part_order, max_part_order;
    int n_windows_kind;      /* 1: tukey(0.5); 2: subdivide_tukey(2); 3: subdivide_tukey(3) */
} preset_t;

/* docs/sonos-pyflac.txt:6926-6934 */
static const preset_t presets[9] = {
    {0, 0, 0, 0, 3, 1}, {1, 1, 0, 0, 3, 1}, {1, 0, 0, 0, 3, 1},
    {0, 0, 6, 0, 4, 1}, {1, 1, 8, 0, 4, 1}, {1, 0, 8, 0, 5, 1},
    {1, 0, 8, 0, 6, 2}, {1, 0, 12, 0, 6, 2}, {1, 0, 12, 0, 6, 3},
};

typedef struct {
    int type;          /* 0 CONSTANT 1 VERBATIM 2 FIXED 3 LPC */
    int order, wasted, precision, shift;
    int32_t coefs[FO_MAX_LPC_ORDER];
    int method, part_order;
    uint32_t params[1 << 8];
    uint32_t bits;     /* libFLAC-style estimate (used for the decision) */
} sf_choice;

typedef struct {
    uint64_t sums[1 << 9];
    uint32_t cur[1 << 8];
    int64_t x[FO_MAX_BLOCK + 1];
    int32_t res[FO_MAX_BLOCK + 1];
    float win[FO_MAX_BLOCK + 1];
    int64_t ch[FO_MAX_CHANNELS][FO_MAX_BLOCK + 1];
    int32_t chres[FO_MAX_CHANNELS][FO_MAX_BLOCK + 1];
    sf_choice choice[FO_MAX_CHANNELS];
    sf_choice cand;
} work_t;

static int ilog2u(uint32_t v) { int r = -1; while (v) { r++; v >>= 1; } return r; }
static int ilog2u64(uint64_t v) { int r = -1; while (v) { r++; v >>= 1; } return r; }

static void window_tukey(float *w, int L, float p) {
    for (int n = 0; n < L; n++) w[n] = 1.0f;
    if (p <= 0.0f) return;
    if (p >= 1.0f) { for (int n = 0; n < L; n++) w[n] = (float)(0.5f - 0.5f * cosf(2.0f * (float)M_PI * n / (L - 1))); return; }
    int Np = (int)(p / 2.0f * L) - 1;
    if (Np > 0) {
        for (int n = 0; n <= Np; n++) {
            w[n] = (float)(0.5f - 0.5f * cosf((float)(M_PI * n / Np)));
            w[L - Np - 1 + n] = (float)(0.5f - 0.5f * cosf((float)(M_PI * (n + Np) / Np)));
        }
    }
}

static void autocorr(const float *d, uint32_t n, uint32_t lags, double *ac) {
    for (uint32_t l = 0; l < lags; l++) ac[l] = 0.0;
    for (uint32_t i = 0; i < n; i++) {
        double x = d[i];
        uint32_t m = (n - i) < lags ? (n - i) : lags;
        for (uint32_t l = 0; l < m; l++) ac[l] += x * (double)d[i + l];
    }
}

static void levinson(const double *ac, uint32_t *max_order, float lp[][FO_MAX_LPC_ORDER], double *error) {
    double lpc[FO_MAX_LPC_ORDER];
    double err = ac[0];
    for (uint32_t i = 0; i < *max_order; i++) {
        double r = -ac[i + 1];
        for (uint32_t j = 0; j < i; j++) r -= lpc[j] * ac[i - j];
        r /= err;
        lpc[i] = r;
        uint32_t j;
        for (j = 0; j < (i >> 1); j++) {
            double tmp = lpc[j];
            lpc[j] += r * lpc[i - 1 - j];
            lpc[i - 1 - j] += r * tmp;
        }
        if (i & 1) lpc[j] += lpc[j] * r;
        err *= (1.0 - r * r);
        for (j = 0; j <= i; j++) lp[i][j] = (float)(-lpc[j]);
        error[i] = err;
        if (err == 0.0) { *max_order = i + 1; return; }
    }
}

static double expected_bits_scaled(double lpc_error, double error_scale) {
    if (lpc_error > 0.0) {
        double bps = 0.5 * log(error_scale * lpc_error) / M_LN2;
        return bps >= 0.0 ? bps : 0.0;
    } else if (lpc_error < 0.0) return 1e32;
    return 0.0;
}

static uint32_t best_lpc_order(const double *err, uint32_t max_order, uint32_t total, uint32_t overhead) {
    double scale = 0.5 / (double)total;
    uint32_t best = 0; double best_bits = (double)(uint32_t)(-1);
    for (uint32_t i = 0, order = 1; i < max_order; i++, order++) {
        double bits = expected_bits_scaled(err[i], scale) * (double)(total - order) + (double)(order * overhead);
        if (bits < best_bits) { best = i; best_bits = bits; }
    }
    return best + 1;
}

static int quantize_coefs(const float *lp, uint32_t order, uint32_t precision, int32_t *q, int *shift) {
    precision--;
    int32_t qmax = 1 << precision, qmin = -qmax; qmax--;
    double cmax = 0.0;
    for (uint32_t i = 0; i < order; i++) { double d = fabs((double)lp[i]); if (d > cmax) cmax = d; }
    if (cmax <= 0.0) return 2;
    int log2cmax; (void)frexp(cmax, &log2cmax); log2cmax--;
    *shift = (int)precision - log2cmax - 1;
    if (*shift > 15) *shift = 15; else if (*shift < -16) return 1;
    if (*shift >= 0) {
        double e = 0.0;
        for (uint32_t i = 0; i < order; i++) {
            e += (double)lp[i] * (double)(1 << *shift);
            long v = lround(e);
            if (v > qmax) v = qmax; else if (v < qmin) v = qmin;
            e -= (double)v; q[i] = (int32_t)v;
        }
    } else {
        int ns = -*shift; double e = 0.0;
        for (uint32_t i = 0; i < order; i++) {
            e += (double)lp[i] / (double)(1 << ns);
            long v = lround(e);
            if (v > qmax) v = qmax; else if (v < qmin) v = qmin;
            e -= (double)v; q[i] = (int32_t)v;
        }
        *shift = 0;
    }
    return 0;
}

static uint32_t max_part_order_for(uint32_t limit, uint32_t blocksize, uint32_t order) {
    uint32_t m = 0, b = blocksize;
    if (b == 0) return 0;
    while (!(b & 1)) { m++; b >>= 1; }
    if (m > 15) m = 15;
    if (m > limit) m = limit;
    while (m > 0 && (blocksize >> m) <= order) m--;
    return m;
}

static uint32_t rice_bits_est(uint32_t k, uint32_t n, uint64_t sum) {
    uint64_t v = 4 + (uint64_t)(1 + k) * n + (k ? (sum >> (k - 1)) : (sum << 1)) - (n >> 1);
    return v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
}

/* find_best_partition_order_ + set_partitioned_rice_ (estimate-driven, no escapes) */
static uint32_t best_partition(work_t *W, const int32_t *res, uint32_t n_res, uint32_t order, uint32_t k_limit,
                               uint32_t min_po, uint32_t max_po, sf_choice *c) {
    uint32_t blocksize = n_res + order;
    max_po = max_part_order_for(max_po, blocksize, order);
    if (min_po > max_po) min_po = max_po;
    uint64_t *sums = W->sums;
    /* sums at max_po */
    {
        uint32_t parts = 1u << max_po, per = blocksize >> max_po, idx = 0;
        for (uint32_t p = 0; p < parts; p++) {
            uint32_t n = per - (p == 0 ? order : 0);
            uint64_t s = 0;
            for (uint32_t j = 0; j < n; j++) { int64_t v = res[idx++]; s += (uint64_t)(v < 0 ? -v : v); }
            sums[p] = s;
        }
    }
    uint32_t best_bits = 0, best_po = 0;
    uint32_t *cur = W->cur;
    uint64_t *level = sums; uint32_t from = 0;
    for (int po = (int)max_po; po >= (int)min_po; po--) {
        uint32_t parts = 1u << po, base = blocksize >> po;
        uint32_t div_base = 0x40000u / base;
        uint32_t bits = 6; int ok = 1;
        for (uint32_t p = 0; p < parts; p++) {
            uint32_t n = base, div = div_base;
            if (p == 0) { if (n <= order) { ok = 0; break; } n -= order; div = 0x40000u / n; }
            uint64_t mean = level[p];
            uint32_t k;
            if (mean < 2 || (((mean - 1) * div) >> 18) == 0) k = 0;
            else k = (uint32_t)ilog2u64(((mean - 1) * div) >> 18) + 1;
            if (k >= k_limit) k = k_limit - 1;
            cur[p] = k;
            uint32_t pb = rice_bits_est(k, n, mean);
            bits = (bits > 0xFFFFFFFFu - pb) ? 0xFFFFFFFFu : bits + pb;
        }
        if (!ok) break;
        if (best_bits == 0 || bits < best_bits) {
            best_bits = bits; best_po = (uint32_t)po;
            memcpy(c->params, cur, parts * sizeof(uint32_t));
        }
        /* merge to next lower order */
        if (po > (int)min_po) {
            uint64_t *next = sums + from + parts;
            for (uint32_t p = 0; p < parts / 2; p++) next[p] = level[2 * p] + level[2 * p + 1];
            from += parts; level = next;
        }
    }
    c->part_order = (int)best_po;
    c->method = 0;
    for (uint32_t p = 0; p < (1u << best_po); p++) if (c->params[p] >= 15) { c->method = 1; break; }
    return best_bits;
}

static uint32_t fixed_best_order(const int64_t *x, uint32_t n, float bits[5]) {
    /* FLAC__fixed_compute_best_predictor[_wide]: data = signal+4, len = blocksize-4 */
    uint64_t e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
    for (uint32_t i = 4; i < n; i++) {
        int64_t a = x[i], b = x[i - 1], c = x[i - 2], d = x[i - 3], e = x[i - 4];
        e0 += (uint64_t)llabs(a);
        e1 += (uint64_t)llabs(a - b);
        e2 += (uint64_t)llabs(a - 2 * b + c);
        e3 += (uint64_t)llabs(a - 3 * b + 3 * c - d);
        e4 += (uint64_t)llabs(a - 4 * b + 6 * c - 4 * d + e);
    }
    uint32_t order;
    uint64_t m1234 = e1 < e2 ? e1 : e2; if (e3 < m1234) m1234 = e3; if (e4 < m1234) m1234 = e4;
    uint64_t m234 = e2 < e3 ? e2 : e3; if (e4 < m234) m234 = e4;
    uint64_t m34 = e3 < e4 ? e3 : e4;
    if (e0 <= m1234) order = 0; else if (e1 <= m234) order = 1; else if (e2 <= m34) order = 2; else if (e3 <= e4) order = 3; else order = 4;
    double len = (double)(n - 4);
    uint64_t es[5] = {e0, e1, e2, e3, e4};
    for (int k = 0; k < 5; k++) bits[k] = (float)(es[k] > 0 ? log(M_LN2 * (double)es[k] / len) / M_LN2 : 0.0);
    return order;
}

typedef struct {
    const preset_t *ps;
    uint32_t blocksize_cfg, bps_stream, qlp_precision;
    float *window;          /* blocksize_cfg floats */
    int exact;              /* 0: libFLAC estimates; (reserved) */
} enc_ctx;

/* Evaluate one subframe the way process_subframe_ does; fills `best` and its residual. */
static void choose_subframe(work_t *W, const enc_ctx *E, const int64_t *x_in, uint32_t n, uint32_t bps_in, sf_choice *best, int32_t *best_res) {
    int64_t *x = W->x; int32_t *res = W->res; float *win = W->win;
#define cand (W->cand)
    /* wasted bits */
    uint64_t orv = 0;
    for (uint32_t i = 0; i < n; i++) orv |= (uint64_t)x_in[i];
    uint32_t w = 0;
    if (orv) { while (!((orv >> w) & 1)) w++; }
    if (w > bps_in) w = bps_in;
    for (uint32_t i = 0; i < n; i++) x[i] = x_in[i] >> w;
    uint32_t bps = bps_in - w;
    const uint32_t k_limit = E->bps_stream > 16 ? 31 : 15;
    memset(best, 0, sizeof *best);
    best->type = 1; best->wasted = (int)w;
    best->bits = 8 + w + n * bps;                /* VERBATIM baseline */
    if (n <= 4) return;
    float fbits[5];
    uint32_t guess = fixed_best_order(x, n, fbits);
    int is_const = 0;
    if (fbits[1] == 0.0f) { is_const = 1; for (uint32_t i = 1; i < n; i++) if (x[i] != x[0]) { is_const = 0; break; } }
    if (is_const) {
        uint32_t b = 8 + w + bps;
        if (b < best->bits) { best->type = 0; best->bits = b; }
        return;
    }
    /* FIXED, guessed order only (no exhaustive search at any preset) */
    if (!(fbits[guess] >= (float)bps)) {
        uint32_t order = guess; int ok = 1;
        for (uint32_t i = order; i < n; i++) {
            int64_t r;
            switch (order) {
                case 0: r = x[i]; break;
                case 1: r = x[i] - x[i - 1]; break;
                case 2: r = x[i] - 2 * x[i - 1] + x[i - 2]; break;
                case 3: r = x[i] - 3 * x[i - 1] + 3 * x[i - 2] - x[i - 3]; break;
                default: r = x[i] - 4 * x[i - 1] + 6 * x[i - 2] - 4 * x[i - 3] + x[i - 4]; break;
            }
            if (r > INT32_MAX || r <= INT32_MIN) { ok = 0; break; }
            res[i - order] = (int32_t)r;
        }
        if (ok) {
            memset(&cand, 0, sizeof cand);
            uint32_t rb = best_partition(W, res, n - order, order, k_limit, (uint32_t)E->ps->min_part_order, (uint32_t)E->ps->max_part_order, &cand);
            uint32_t b = 8 + w + order * bps + rb;
            if (b < best->bits) {
                cand.type = 2; cand.order = (int)order; cand.wasted = (int)w; cand.bits = b;
                *best = cand; memcpy(best_res, res, (n - order) * sizeof(int32_t));
            }
        }
    }
    /* LPC */
    uint32_t max_lpc = (uint32_t)E->ps->max_lpc_order;
    if (max_lpc >= n) max_lpc = n - 1;
    if (max_lpc == 0) return;
    int kinds = E->ps->n_windows_kind;
    /* enumerate windows: subdivide_tukey(parts): b=1 full; b=2: two halves; b>=3: partial c/b and punchout */
    double ac_root[FO_MAX_LPC_ORDER + 1], ac[FO_MAX_LPC_ORDER + 1];
    for (int b = 1; b <= kinds; b++) {
        int ncs = (b == 1) ? 1 : (b == 2 ? 2 : 2 * b);
        for (int ci = 0; ci < ncs; ci++) {
            int c = (b == 2) ? 2 * ci : ci;        /* b==2: only even c (halves) */
            if (b == 1) {
                for (uint32_t i = 0; i < n; i++) win[i] = (float)x[i] * E->window[i];
                autocorr(win, n, max_lpc + 1, ac);
                memcpy(ac_root, ac, sizeof ac);
            } else {
                if (n / (uint32_t)b <= FO_MAX_LPC_ORDER) continue;
                if (!(c & 1)) {
                    uint32_t part = n / (uint32_t)b / 2, shift = ((uint32_t)(c / 2) * n) / (uint32_t)b;
                    uint32_t plen = n / (uint32_t)b;
                    /* FLAC__lpc_window_data_partial */
                    if (part + shift < n) {
                        uint32_t i, j;
                        for (i = 0; i < part; i++) win[i] = (float)x[shift + i] * E->window[i];
                        uint32_t lim = n - part - shift; if (i > lim) i = lim;
                        for (j = n - part; j < n; i++, j++) win[i] = (float)x[shift + i] * E->window[j];
                        if (i < n) win[i] = 0.0f;
                    }
                    autocorr(win, plen, max_lpc + 1, ac);
                } else {
                    for (uint32_t l = 0; l < max_lpc; l++) ac[l] = ac_root[l] - ac[l];
                }
            }
            if (ac[0] == 0.0) continue;
            float lp[FO_MAX_LPC_ORDER][FO_MAX_LPC_ORDER]; double err[FO_MAX_LPC_ORDER];
            uint32_t mo = max_lpc;
            levinson(ac, &mo, lp, err);
            uint32_t order = best_lpc_order(err, mo, n, bps + E->qlp_precision);
            double ebits = expected_bits_scaled(err[order - 1], 0.5 / (double)(n - order));
            if (ebits >= (double)bps) continue;
            uint32_t prec = E->qlp_precision;
            if (bps <= 17) { uint32_t lim = 32 - bps - (uint32_t)ilog2u(order); if (lim < prec) prec = lim; }
            int32_t q[FO_MAX_LPC_ORDER]; int shift;
            if (quantize_coefs(lp[order - 1], order, prec, q, &shift)) continue;
            int ok = 1;
            for (uint32_t i = order; i < n; i++) {
                int64_t s = 0;
                for (uint32_t j = 0; j < order; j++) s += (int64_t)q[j] * x[i - 1 - j];
                int64_t r = x[i] - (s >> shift);
                if (r > INT32_MAX || r <= INT32_MIN) { ok = 0; break; }
                res[i - order] = (int32_t)r;
            }
            if (!ok) continue;
            memset(&cand, 0, sizeof cand);
            uint32_t rb = best_partition(W, res, n - order, order, k_limit, (uint32_t)E->ps->min_part_order, (uint32_t)E->ps->max_part_order, &cand);
            uint32_t bt = 8 + w + 4 + 5 + order * (prec + bps) + rb;
            if (bt < best->bits) {
                cand.type = 3; cand.order = (int)order; cand.wasted = (int)w; cand.precision = (int)prec; cand.shift = shift;
                memcpy(cand.coefs, q, order * sizeof(int32_t)); cand.bits = bt;
                *best = cand; memcpy(best_res, res, (n - order) * sizeof(int32_t));
            }
        }
    }
}

#undef cand

static void write_subframe(bw_t *w, const sf_choice *c, const int64_t *x_in, const int32_t *res, uint32_t n, uint32_t bps_in) {
    uint32_t bps = bps_in - (uint32_t)c->wasted;
    uint32_t typecode = c->type == 0 ? 0 : c->type == 1 ? 1 : c->type == 2 ? (8u | (uint32_t)c->order) : (32u | (uint32_t)(c->order - 1));
    bw_bits(w, (typecode << 1) | (c->wasted ? 1u : 0u), 8);
    if (c->wasted) bw_unary(w, (uint32_t)c->wasted - 1);
    uint64_t mask = bps >= 64 ? ~0ull : ((1ull << bps) - 1ull);
    if (c->type == 0) { bw_bits(w, (uint64_t)(x_in[0] >> c->wasted) & mask, (int)bps); return; }
    if (c->type == 1) { for (uint32_t i = 0; i < n; i++) bw_bits(w, (uint64_t)(x_in[i] >> c->wasted) & mask, (int)bps); return; }
    for (int i = 0; i < c->order; i++) bw_bits(w, (uint64_t)(x_in[i] >> c->wasted) & mask, (int)bps);
    if (c->type == 3) {
        bw_bits(w, (uint64_t)(c->precision - 1), 4);
        bw_bits(w, (uint64_t)c->shift & 31u, 5);
        for (int i = 0; i < c->order; i++) bw_bits(w, (uint64_t)(int64_t)c->coefs[i] & ((1ull << c->precision) - 1ull), c->precision);
    }
    bw_bits(w, (uint64_t)c->method, 2);
    bw_bits(w, (uint64_t)c->part_order, 4);
    uint32_t parts = 1u << c->part_order, idx = 0;
    for (uint32_t p = 0; p < parts; p++) {
        uint32_t cnt = (n >> c->part_order) - (p == 0 ? (uint32_t)c->order : 0);
        bw_bits(w, c->params[p], c->method ? 5 : 4);
        for (uint32_t j = 0; j < cnt; j++) bw_rice(w, res[idx++], c->params[p]);
    }
}

static uint32_t exact_subframe_bits(const sf_choice *c, const int32_t *res, uint32_t n, uint32_t bps_in) {
    uint32_t bps = bps_in - (uint32_t)c->wasted;
    uint32_t b = 8 + (uint32_t)c->wasted;
    if (c->type == 0) return b + bps;
    if (c->type == 1) return b + n * bps;
    b += (uint32_t)c->order * bps;
    if (c->type == 3) b += 9 + (uint32_t)(c->order * c->precision);
    b += 6;
    uint32_t parts = 1u << c->part_order, idx = 0;
    for (uint32_t p = 0; p < parts; p++) {
        uint32_t cnt = (n >> c->part_order) - (p == 0 ? (uint32_t)c->order : 0);
        uint32_t k = c->params[p];
        b += c->method ? 5 : 4;
        for (uint32_t j = 0; j < cnt; j++) { int32_t v = res[idx++]; uint32_t u = ((uint32_t)v << 1) ^ (uint32_t)(v >> 31); b += (u >> k) + 1 + k; }
    }
    return b;
}

static int utf8_put(uint8_t *p, uint64_t v) {
    if (v < 0x80) { p[0] = (uint8_t)v; return 1; }
    int n = v < 0x800 ? 2 : v < 0x10000 ? 3 : v < 0x200000 ? 4 : v < 0x4000000 ? 5 : v < 0x80000000ull ? 6 : 7;
    static const uint8_t lead[8] = {0, 0, 0xC0, 0xE0, 0xF0, 0xF8, 0xFC, 0xFE};
    for (int i = n - 1; i > 0; i--) { p[i] = (uint8_t)(0x80 | (v & 0x3F)); v >>= 6; }
    p[0] = (uint8_t)(lead[n] | v);
    return n;
}

static size_t write_frame_header(uint8_t *p, uint32_t blocksize, uint32_t sample_rate, uint32_t ch_assign, uint32_t bps, uint64_t number) {
    uint32_t bsc, src, bpc; int bs_hint = 0, sr_hint = 0;
    switch (blocksize) {
        case 192: bsc = 1; break; case 576: bsc = 2; break; case 1152: bsc = 3; break; case 2304: bsc = 4; break; case 4608: bsc = 5; break;
        case 256: bsc = 8; break; case 512: bsc = 9; break; case 1024: bsc = 10; break; case 2048: bsc = 11; break; case 4096: bsc = 12; break;
        case 8192: bsc = 13; break; case 16384: bsc = 14; break; case 32768: bsc = 15; break;
        default: if (blocksize <= 256) { bsc = 6; bs_hint = 1; } else { bsc = 7; bs_hint = 2; }
    }
    switch (sample_rate) {
        case 88200: src = 1; break; case 176400: src = 2; break; case 192000: src = 3; break; case 8000: src = 4; break;
        case 16000: src = 5; break; case 22050: src = 6; break; case 24000: src = 7; break; case 32000: src = 8; break;
        case 44100: src = 9; break; case 48000: src = 10; break; case 96000: src = 11; break;
        default:
            if (sample_rate <= 255000 && sample_rate % 1000 == 0) { src = 12; sr_hint = 1; }
            else if (sample_rate <= 655350 && sample_rate % 10 == 0) { src = 14; sr_hint = 3; }
            else if (sample_rate <= 0xFFFF) { src = 13; sr_hint = 2; }
            else src = 0;
    }
    switch (bps) { case 8: bpc = 1; break; case 12: bpc = 2; break; case 16: bpc = 4; break; case 20: bpc = 5; break; case 24: bpc = 6; break; case 32: bpc = 7; break; default: bpc = 0; }
    size_t i = 0;
    p[i++] = 0xFF; p[i++] = 0xF8;
    p[i++] = (uint8_t)((bsc << 4) | src);
    p[i++] = (uint8_t)((ch_assign << 4) | (bpc << 1));
    i += (size_t)utf8_put(p + i, number);
    if (bs_hint == 1) p[i++] = (uint8_t)(blocksize - 1);
    else if (bs_hint == 2) { p[i++] = (uint8_t)((blocksize - 1) >> 8); p[i++] = (uint8_t)(blocksize - 1); }
    if (sr_hint == 1) p[i++] = (uint8_t)(sample_rate / 1000);
    else if (sr_hint == 2) { p[i++] = (uint8_t)(sample_rate >> 8); p[i++] = (uint8_t)sample_rate; }
    else if (sr_hint == 3) { p[i++] = (uint8_t)((sample_rate / 10) >> 8); p[i++] = (uint8_t)(sample_rate / 10); }
    p[i] = fo_crc8(p, i); i++;
    return i;
}

static const char FO_VENDOR[] = "flac-raster-b200 oracle (libFLAC 1.4.3 procedure restated)";

/*
 * Encode interleaved int32 samples (the buffer pyflac passes to
 * FLAC__stream_encoder_process_interleaved, docs/sonos-pyflac.txt:1994-1997)
 * into a complete FLAC stream: "fLaC" + STREAMINFO + VORBIS_COMMENT(vendor
 * only) + frames, exactly the chunks the reference's write callback receives
 * (converter.py:392-400).  With finalize==0 STREAMINFO keeps total_samples /
 * framesizes / md5 at zero, as in every reference-made file (no seek callback,
 * converter.py:139-144).  Returns bytes written, 0 on overflow/error.
 * frame_sizes (optional) receives per-frame byte counts.
 */
size_t fo_encode_stream(const int32_t *samples, uint64_t n_samples, uint32_t channels, uint32_t bps,
                        uint32_t sample_rate, uint32_t level, uint32_t blocksize, int finalize,
                        const char *vendor, uint8_t *out, size_t cap,
                        uint32_t *frame_sizes, size_t frame_cap, fo_subframe_desc *descs, size_t desc_cap) {
    if (level > 8 || channels < 1 || channels > 8 || blocksize < 16 || blocksize > FO_MAX_BLOCK) return 0;
    if (!vendor) vendor = FO_VENDOR;
    size_t vlen = strlen(vendor);
    size_t hdr = 4 + 4 + 34 + 4 + 4 + vlen + 4;
    if (cap < hdr) return 0;
    uint8_t *p = out;
    memcpy(p, "fLaC", 4); p += 4;
    uint8_t *si = p;
    memset(si, 0, 38);
    si[0] = 0; si[3] = 34;
    uint8_t *s = si + 4;
    s[0] = (uint8_t)(blocksize >> 8); s[1] = (uint8_t)blocksize; s[2] = s[0]; s[3] = s[1];
    s[10] = (uint8_t)(sample_rate >> 12); s[11] = (uint8_t)(sample_rate >> 4);
    s[12] = (uint8_t)(((sample_rate & 15) << 4) | ((channels - 1) << 1) | (((bps - 1) >> 4) & 1));
    s[13] = (uint8_t)(((bps - 1) & 15) << 4);
    p += 38;
    p[0] = 0x84; p[1] = (uint8_t)((8 + vlen) >> 16); p[2] = (uint8_t)((8 + vlen) >> 8); p[3] = (uint8_t)(8 + vlen); p += 4;
    p[0] = (uint8_t)vlen; p[1] = (uint8_t)(vlen >> 8); p[2] = (uint8_t)(vlen >> 16); p[3] = (uint8_t)(vlen >> 24); p += 4;
    memcpy(p, vendor, vlen); p += vlen;
    memset(p, 0, 4); p += 4;

    enc_ctx E; E.ps = &presets[level]; E.blocksize_cfg = blocksize; E.bps_stream = bps; E.exact = 0;
    if (bps < 16) { E.qlp_precision = 2 + bps / 2; if (E.qlp_precision < 5) E.qlp_precision = 5; }
    else if (bps == 16) E.qlp_precision = blocksize <= 192 ? 7 : blocksize <= 384 ? 8 : blocksize <= 576 ? 9 : blocksize <= 1152 ? 10 : blocksize <= 2304 ? 11 : blocksize <= 4608 ? 12 : 13;
    else E.qlp_precision = blocksize <= 384 ? 13 : blocksize <= 1152 ? 14 : 15;
    E.window = (float *)malloc(sizeof(float) * blocksize);
    window_tukey(E.window, (int)blocksize, 0.5f / (float)E.ps->n_windows_kind);

    work_t *W = malloc(sizeof(work_t));
    if (!W) { free(E.window); return 0; }
    int64_t (*ch)[FO_MAX_BLOCK + 1] = W->ch; int32_t (*res)[FO_MAX_BLOCK + 1] = W->chres; sf_choice *choice = W->choice;
    uint64_t pos = 0, frame_no = 0; size_t nd = 0;
    /* loose mid/side state (FLAC__stream_encoder: loose_mid_side_stereo_frames = sample_rate * 0.4 / blocksize) */
    uint32_t loose_frames = (uint32_t)((double)sample_rate * 0.4 / (double)blocksize + 0.5), loose_count = 0, last_assign = 1;
    if (loose_frames == 0) loose_frames = 1;
    uint32_t minf = 0xFFFFFFFFu, maxf = 0;
    while (pos < n_samples) {
        uint32_t n = (n_samples - pos) < blocksize ? (uint32_t)(n_samples - pos) : blocksize;
        for (uint32_t c = 0; c < channels; c++)
            for (uint32_t i = 0; i < n; i++) ch[c][i] = samples[(pos + i) * channels + c];
        /* Stereo decorrelation for exactly two channels (libFLAC process_subframes_ / process_frame_, presets with
         * do_mid_side): mid = (L+R)>>1, side = L-R (one more bit per sample); all four subframes are evaluated
         * and the cheapest of independent / left-side / right-side / mid-side by the estimated bits wins, ties
         * going to the earlier one.  "Loose" presets (levels 1 and 4) take that decision only every
         * loose_frames frames and otherwise keep independent or switch to mid-side.  [upstream, unverified:
         * restated from libFLAC 1.4.3; the reference holds no 2-channel golden]  finalize bit 1 disables it. */
        const int do_ms = (channels == 2) && E.ps->do_mid_side && !(finalize & 2);
        uint32_t ch_assign = channels - 1;
        uint32_t order_ch[2] = {0, 1};       /* which of L,R,M,S go into the frame */
        uint32_t nsub = channels;
        if (do_ms) {
            for (uint32_t i = 0; i < n; i++) { ch[2][i] = (ch[0][i] + ch[1][i]) >> 1; ch[3][i] = ch[0][i] - ch[1][i]; }
            int eval_indep = 1, eval_ms = 1;
            if (E.ps->loose_mid_side && loose_count > 0) { eval_indep = (last_assign == 1); eval_ms = !eval_indep; }
            if (eval_indep) { choose_subframe(W, &E, ch[0], n, bps, &choice[0], res[0]); choose_subframe(W, &E, ch[1], n, bps, &choice[1], res[1]); }
            if (eval_ms) { choose_subframe(W, &E, ch[2], n, bps, &choice[2], res[2]); choose_subframe(W, &E, ch[3], n, bps + 1, &choice[3], res[3]); }
            if (E.ps->loose_mid_side && loose_count > 0) ch_assign = eval_indep ? 1 : 10;
            else {
                uint64_t b[4] = { (uint64_t)choice[0].bits + choice[1].bits, (uint64_t)choice[0].bits + choice[3].bits,
                                  (uint64_t)choice[1].bits + choice[3].bits, (uint64_t)choice[2].bits + choice[3].bits };
                uint32_t best = 0;
                for (uint32_t a = 1; a < 4; a++) if (b[a] < b[best]) best = a;
                ch_assign = best == 0 ? 1 : 7 + best;          /* 8 left/side, 9 side/right, 10 mid/side */
            }
            if (E.ps->loose_mid_side) { loose_count++; if (loose_count >= loose_frames) loose_count = 0; }
            last_assign = ch_assign;
            if (ch_assign == 8) { order_ch[0] = 0; order_ch[1] = 3; }
            else if (ch_assign == 9) { order_ch[0] = 3; order_ch[1] = 1; }
            else if (ch_assign == 10) { order_ch[0] = 2; order_ch[1] = 3; }
        }
        size_t need = 16 + (size_t)channels * ((size_t)n * (bps + 2) / 8 + 16) + 2;
        if ((size_t)(p - out) + need > cap) { free(E.window); free(W); return 0; }
        uint8_t *f0 = p;
        size_t hb = write_frame_header(p, n, sample_rate, ch_assign, bps, frame_no);
        bw_t w = { f0, need, hb * 8, 0 };
        for (uint32_t k = 0; k < nsub; k++) {
            const uint32_t c = do_ms ? order_ch[k] : k;
            const uint32_t sbps = bps + ((do_ms && c == 3) ? 1u : 0u);
            if (!do_ms) choose_subframe(W, &E, ch[c], n, bps, &choice[c], res[c]);
            size_t b0 = w.bitpos;
            write_subframe(&w, &choice[c], ch[c], res[c], n, sbps);
            if (descs && nd < desc_cap) {
                fo_subframe_desc *d = &descs[nd++];
                memset(d, 0, sizeof *d);
                d->frame = (uint32_t)frame_no; d->channel = k; d->type = (uint32_t)choice[c].type; d->order = (uint32_t)choice[c].order;
                d->wasted = (uint32_t)choice[c].wasted; d->precision = (uint32_t)choice[c].precision; d->shift = choice[c].shift;
                memcpy(d->coefs, choice[c].coefs, sizeof d->coefs); d->method = (uint32_t)choice[c].method; d->partition_order = (uint32_t)choice[c].part_order;
                for (uint32_t q = 0; q < FO_DESC_PARAMS && q < (1u << choice[c].part_order); q++) d->params[q] = choice[c].params[q];
                d->nbits = (uint32_t)(w.bitpos - b0); d->blocksize = n; d->ch_assign = ch_assign; d->frame_offset = (uint32_t)(f0 - out);
                (void)exact_subframe_bits;
            }
        }
        if (w.bitpos & 7) bw_bits(&w, 0, 8 - (int)(w.bitpos & 7));
        if (w.err) { free(E.window); free(W); return 0; }
        size_t fb = w.bitpos >> 3;
        uint16_t crc = fo_crc16(f0, fb);
        f0[fb] = (uint8_t)(crc >> 8); f0[fb + 1] = (uint8_t)crc; fb += 2;
        if (frame_sizes && frame_no < frame_cap) frame_sizes[frame_no] = (uint32_t)fb;
        if (fb < minf) minf = (uint32_t)fb;
        if (fb > maxf) maxf = (uint32_t)fb;
        p += fb; pos += n; frame_no++;
    }
    free(E.window); free(W);
    if ((finalize & 1) && frame_no) {
        s[4] = (uint8_t)(minf >> 16); s[5] = (uint8_t)(minf >> 8); s[6] = (uint8_t)minf;
        s[7] = (uint8_t)(maxf >> 16); s[8] = (uint8_t)(maxf >> 8); s[9] = (uint8_t)maxf;
        s[13] = (uint8_t)((s[13] & 0xF0) | ((n_samples >> 32) & 15));
        s[14] = (uint8_t)(n_samples >> 24); s[15] = (uint8_t)(n_samples >> 16); s[16] = (uint8_t)(n_samples >> 8); s[17] = (uint8_t)n_samples;
    }
    return (size_t)(p - out);
}

size_t fo_sizeof_info(void) { return sizeof(fo_stream_info); }
size_t fo_sizeof_desc(void) { return sizeof(fo_subframe_desc); }
