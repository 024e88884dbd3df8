// frb_hostparse.cuh -- host-side (CPU, no CUDA) walk over the metadata blocks of many tile files at once.
//
// A bbox query over a streaming container (cli.py:297-315 looped; README get_tiles_by_bbox) fetches thousands of small
// standalone FLAC files.  Their frames are decoded in one GPU batch, but round 1 parsed every file's STREAMINFO and
// VORBIS_COMMENT tags in Python: 0.27 s for 4096 tiles against 5 ms of GPU decode.  These two calls do the same walk
// in C: what the decode needs (frame offset, STREAMINFO, the GEOSPATIAL_* numbers the reference parses at
// converter.py:356-377, the seek index block) comes back as one array of structs.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace frb {
static inline bool tag_is(const uint8_t *e, uint32_t elen, const char *key, const uint8_t **val, uint32_t *vlen) {
    const size_t kl = strlen(key);
    if (elen <= kl || e[kl] != '=') return false;
    for (size_t i = 0; i < kl; i++) {
        uint8_t c = e[i];
        if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
        if (c != (uint8_t)key[i]) return false;
    }
    *val = e + kl + 1; *vlen = elen - (uint32_t)kl - 1;
    return true;
}
static inline double tag_double(const uint8_t *v, uint32_t n, bool *ok) {
    char buf[64];
    if (n == 0 || n >= sizeof buf) { *ok = false; return NAN; }
    memcpy(buf, v, n); buf[n] = 0;
    char *end = nullptr;
    const double d = strtod(buf, &end);
    *ok = end != buf && *end == 0;
    return d;
}
static inline int dtype_code(const uint8_t *v, uint32_t n) {
    static const char *names[8] = {"uint8", "int8", "uint16", "int16", "uint32", "int32", "float32", "float64"};
    for (int i = 0; i < 8; i++) if (strlen(names[i]) == n && memcmp(names[i], v, n) == 0) return i;
    return -1;
}
}  // namespace frb

extern "C" int frb_parse_tile_headers(const uint8_t *base, const uint64_t *offsets, const uint64_t *sizes, uint32_t n,
                                      frb_tile_header *out) {
    using namespace frb;
    if (!base || !offsets || !sizes || !out) return FRB_ERR_INVALID_ARG;
    int bad = 0;
    for (uint32_t t = 0; t < n; t++) {
        frb_tile_header &H = out[t];
        memset(&H, 0, sizeof H);
        H.data_min = H.data_max = H.nodata = NAN;
        H.dtype = -1;
        const uint8_t *b = base + offsets[t];
        const uint64_t len = sizes[t];
        if (len < 42 || memcmp(b, "fLaC", 4) != 0) { bad++; continue; }
        uint64_t pos = 4;
        bool last = false, have_si = false, ok = true;
        while (!last) {
            if (pos + 4 > len) { ok = false; break; }
            last = (b[pos] & 0x80) != 0;
            const uint32_t type = b[pos] & 0x7F;
            const uint64_t blen = ((uint64_t)b[pos + 1] << 16) | ((uint64_t)b[pos + 2] << 8) | b[pos + 3];
            pos += 4;
            if (pos + blen > len || type == 127) { ok = false; break; }
            const uint8_t *s = b + pos;
            if (type == 0 && blen >= 34) {
                H.min_blocksize = ((uint32_t)s[0] << 8) | s[1];
                H.max_blocksize = ((uint32_t)s[2] << 8) | s[3];
                H.sample_rate = ((uint32_t)s[10] << 12) | ((uint32_t)s[11] << 4) | (s[12] >> 4);
                H.channels = ((s[12] >> 1) & 7) + 1;
                H.bps = (((uint32_t)(s[12] & 1) << 4) | (s[13] >> 4)) + 1;
                H.total_samples = ((uint64_t)(s[13] & 15) << 32) | ((uint64_t)s[14] << 24) | ((uint64_t)s[15] << 16) | ((uint64_t)s[16] << 8) | s[17];
                have_si = true;
            } else if (type == 4 && blen >= 8) {
                // VORBIS_COMMENT: little-endian lengths (RFC 9639 8.6); first value of a repeated key wins, like the
                // reference's tags[field][0] (converter.py:358)
                uint64_t q = 0;
                auto rd = [&](uint32_t *v) { if (q + 4 > blen) return false; *v = (uint32_t)s[q] | ((uint32_t)s[q + 1] << 8) | ((uint32_t)s[q + 2] << 16) | ((uint32_t)s[q + 3] << 24); q += 4; return true; };
                uint32_t vl = 0, cnt = 0;
                if (rd(&vl) && q + vl <= blen) {
                    q += vl;
                    if (rd(&cnt)) {
                        uint32_t seen = 0;
                        for (uint32_t i = 0; i < cnt; i++) {
                            uint32_t el = 0;
                            if (!rd(&el) || q + el > blen) break;
                            const uint8_t *e = s + q, *v = nullptr;
                            uint32_t vn = 0;
                            bool okd = false;
                            if (!(seen & 1) && tag_is(e, el, "GEOSPATIAL_CRS", &v, &vn)) { seen |= 1; H.flags |= 1u; }
                            else if (!(seen & 2) && tag_is(e, el, "GEOSPATIAL_WIDTH", &v, &vn)) { seen |= 2; H.width = (uint32_t)tag_double(v, vn, &okd); }
                            else if (!(seen & 4) && tag_is(e, el, "GEOSPATIAL_HEIGHT", &v, &vn)) { seen |= 4; H.height = (uint32_t)tag_double(v, vn, &okd); }
                            else if (!(seen & 8) && tag_is(e, el, "GEOSPATIAL_COUNT", &v, &vn)) { seen |= 8; H.count = (uint32_t)tag_double(v, vn, &okd); }
                            else if (!(seen & 16) && tag_is(e, el, "GEOSPATIAL_DTYPE", &v, &vn)) { seen |= 16; H.dtype = dtype_code(v, vn); }
                            else if (!(seen & 32) && tag_is(e, el, "GEOSPATIAL_DATA_MIN", &v, &vn)) { seen |= 32; const double d = tag_double(v, vn, &okd); H.data_min = okd ? d : 0.0; }
                            else if (!(seen & 64) && tag_is(e, el, "GEOSPATIAL_DATA_MAX", &v, &vn)) { seen |= 64; const double d = tag_double(v, vn, &okd); H.data_max = okd ? d : 0.0; }
                            else if (!(seen & 128) && tag_is(e, el, "GEOSPATIAL_NODATA", &v, &vn)) {
                                seen |= 128;
                                const double d = tag_double(v, vn, &okd);                 // "None" / "" -> absent
                                if (okd) { H.nodata = d; H.flags |= 2u; }
                            }
                            q += el;
                        }
                    }
                }
            } else if (type == 2 && blen >= 4 && memcmp(s, "frbI", 4) == 0) {
                H.flags |= 4u;
                H.index_offset = (uint32_t)(pos + 4);
                H.index_len = (uint32_t)(blen - 4);
            }
            pos += blen;
        }
        if (!ok || !have_si || pos > 0xFFFFFFFFull) { memset(&H, 0, sizeof H); H.dtype = -1; bad++; continue; }
        H.first_frame_offset = (uint32_t)pos;
    }
    return bad ? FRB_ERR_BAD_STREAM : FRB_OK;
}

extern "C" int frb_gather_seek_index(const uint8_t *base, const uint64_t *offsets, const frb_tile_header *hdrs, uint32_t n,
                                     uint32_t channels, uint32_t blocksize, const uint32_t *frames_per_tile,
                                     uint32_t *frame_bytes_out, uint32_t *sub_bitoff_out) {
    if (!base || !offsets || !hdrs || !frames_per_tile || !frame_bytes_out || (channels > 1 && !sub_bitoff_out)) return FRB_ERR_INVALID_ARG;
    uint64_t f0 = 0;
    for (uint32_t t = 0; t < n; t++) {
        const frb_tile_header &H = hdrs[t];
        const uint32_t nf = frames_per_tile[t];
        if (!(H.flags & 4u)) return FRB_ERR_BAD_STREAM;
        const uint8_t *d = base + offsets[t] + H.index_offset;
        const uint64_t need = 12 + 4ull * nf + (channels > 1 ? 4ull * nf * channels : 0);
        if (H.index_len != need) return FRB_ERR_BAD_STREAM;
        uint32_t hb, hn;
        memcpy(&hb, d + 4, 4); memcpy(&hn, d + 8, 4);
        if (d[0] != 1 || d[1] != channels || hb != blocksize || hn != nf) return FRB_ERR_BAD_STREAM;
        memcpy(frame_bytes_out + f0, d + 12, 4ull * nf);
        if (channels > 1) memcpy(sub_bitoff_out + f0 * channels, d + 12 + 4ull * nf, 4ull * nf * channels);
        f0 += nf;
    }
    return FRB_OK;
}
