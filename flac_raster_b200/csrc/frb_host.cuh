// frb_host.cuh -- host-buffer entry points (the e2e path) and the libFLAC-shaped handle API.
//
// frb_host_encode / frb_host_decode are what the Python StreamEncoder /
// FileDecoder shims call through ctypes with numpy (host) buffers: H2D copy,
// kernels, D2H copy, all inside the call.  The frb_stream_encoder_* functions
// mirror the FLAC__stream_encoder_* calls pyflac's cffi layer makes
// (docs/sonos-pyflac.txt:2200-2212, :1994-1997, :2003-2014).
#pragma once
#include "frb_common.cuh"
#include <vector>
#include <string>
#include <string.h>

namespace frb {

// grow-only device / pinned scratch, one set per host thread
struct Scratch {
    void *ptr = nullptr; size_t cap = 0; bool pinned = false;
    int reserve(size_t bytes) {
        if (bytes <= cap) return FRB_OK;
        if (ptr) { if (pinned) cudaFreeHost(ptr); else cudaFree(ptr); ptr = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = pinned ? cudaMallocHost(&ptr, want) : cudaMalloc(&ptr, want);
        if (e != cudaSuccess) return cuda_fail(e, pinned ? "cudaMallocHost" : "cudaMalloc");
        cap = want;
        return FRB_OK;
    }
};
struct HostCtx {
    Scratch d_in, d_audio, d_ws, d_out, d_misc;
    Scratch h_pin;
    cudaStream_t stream = nullptr;
    HostCtx() { h_pin.pinned = true; }
    int init() {
        if (!stream) FRB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        return FRB_OK;
    }
};
static thread_local HostCtx t_ctx;

__global__ void __launch_bounds__(256)
k_deinterleave(const int32_t *__restrict__ in, int32_t *__restrict__ out, uint64_t n, uint32_t channels) {
    const uint64_t total = n * channels;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = i / channels; uint32_t c = (uint32_t)(i - s * channels);
        out[(uint64_t)c * n + s] = in[i];
    }
}
__global__ void __launch_bounds__(256)
k_interleave(const int32_t *__restrict__ in, int32_t *__restrict__ out, uint64_t n, uint32_t channels) {
    const uint64_t total = n * channels;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = i / channels; uint32_t c = (uint32_t)(i - s * channels);
        out[i] = in[(uint64_t)c * n + s];
    }
}

}  // namespace frb

extern "C" int frb_host_encode(const int32_t *interleaved, uint64_t n_samples, uint32_t channels,
                               uint32_t bps, uint32_t sample_rate, uint32_t level, uint32_t blocksize,
                               uint64_t first_frame_number,
                               uint8_t *out, size_t out_capacity, size_t *out_bytes,
                               uint32_t *frame_bytes, size_t frame_capacity, size_t *n_frames) {
    using namespace frb;
    if (!interleaved || !out || !out_bytes || n_samples == 0) return FRB_ERR_INVALID_ARG;
    if (first_frame_number != 0) return FRB_ERR_UNSUPPORTED;   // segments restart numbering per call (see handle API)
    frb_encode_params p = {1, channels, bps, blocksize, level, 0};
    if (!enc_params_ok(&p)) return FRB_ERR_INVALID_ARG;
    HostCtx &C = t_ctx;
    int rc = C.init();
    if (rc) return rc;
    const uint64_t total = n_samples * channels;
    const uint64_t frames = (n_samples + blocksize - 1) / blocksize;
    size_t ws_bytes = 0;
    if ((rc = frb_encode_workspace_size(&p, frames, &ws_bytes))) return rc;
    if ((rc = C.d_in.reserve(total * 4))) return rc;
    if ((rc = C.d_audio.reserve(total * 4))) return rc;
    if ((rc = C.d_ws.reserve(ws_bytes))) return rc;
    cudaStream_t s = C.stream;
    FRB_CUDA(cudaMemcpyAsync(C.d_in.ptr, interleaved, total * 4, cudaMemcpyHostToDevice, s));
    if (channels == 1) {
        FRB_CUDA(cudaMemcpyAsync(C.d_audio.ptr, C.d_in.ptr, total * 4, cudaMemcpyDeviceToDevice, s));
    } else {
        k_deinterleave<<<grid_for(total, 256 * 8, kNumSMs * 16), 256, 0, s>>>((const int32_t *)C.d_in.ptr, (int32_t *)C.d_audio.ptr, n_samples, channels);
        FRB_LAUNCH_CHECK("k_deinterleave");
    }
    uint64_t hn = n_samples; uint32_t hr = sample_rate; int64_t hb = 0; uint64_t total_bytes = 0;
    if ((rc = frb_encode_analyse(&p, (const int32_t *)C.d_audio.ptr, &hn, &hr, &hb, C.d_ws.ptr, C.d_ws.cap, nullptr, &total_bytes, s))) return rc;
    *out_bytes = (size_t)total_bytes;
    if (n_frames) *n_frames = (size_t)frames;
    if (total_bytes > out_capacity) return FRB_ERR_OVERFLOW;
    if ((rc = C.d_out.reserve(total_bytes + 16))) return rc;
    if ((rc = C.d_misc.reserve(frames * 4 + 64))) return rc;
    uint64_t off0 = 0;
    if ((rc = frb_encode_emit(&p, C.d_ws.ptr, C.d_ws.cap, &off0, (uint8_t *)C.d_out.ptr, C.d_out.cap, (uint32_t *)C.d_misc.ptr, s))) return rc;
    FRB_CUDA(cudaMemcpyAsync(out, C.d_out.ptr, total_bytes, cudaMemcpyDeviceToHost, s));
    if (frame_bytes) {
        size_t nf = frames < frame_capacity ? (size_t)frames : frame_capacity;
        FRB_CUDA(cudaMemcpyAsync(frame_bytes, C.d_misc.ptr, nf * 4, cudaMemcpyDeviceToHost, s));
    }
    FRB_CUDA(cudaStreamSynchronize(s));
    return FRB_OK;
}

extern "C" int frb_host_decode(const uint8_t *frames, size_t n_bytes, uint32_t channels, uint32_t bps,
                               uint32_t blocksize, uint32_t sample_rate, uint64_t n_samples_hint,
                               int32_t *interleaved_out, size_t out_capacity_samples,
                               uint64_t *n_samples_out) {
    using namespace frb;
    if (!frames || !n_bytes || !n_samples_out) return FRB_ERR_INVALID_ARG;
    HostCtx &C = t_ctx;
    int rc = C.init();
    if (rc) return rc;
    cudaStream_t s = C.stream;
    if ((rc = C.d_in.reserve(n_bytes + 64))) return rc;
    FRB_CUDA(cudaMemsetAsync((uint8_t *)C.d_in.ptr + (n_bytes & ~(size_t)3), 0, 32, s));
    FRB_CUDA(cudaMemcpyAsync(C.d_in.ptr, frames, n_bytes, cudaMemcpyHostToDevice, s));
    uint64_t n_samples = n_samples_hint, n_frames = 0;
    if (n_samples == 0) {
        if ((rc = frb_probe_stream((const uint8_t *)C.d_in.ptr, 0, n_bytes, channels, bps, blocksize, sample_rate, &n_frames, &n_samples, s))) return rc;
    }
    n_frames = (n_samples + blocksize - 1) / blocksize;
    *n_samples_out = n_samples;
    if (!interleaved_out) return FRB_OK;                      // size query
    const uint64_t total = n_samples * channels;
    if (total > out_capacity_samples) return FRB_ERR_OVERFLOW;
    frb_decode_params p = {1, channels, bps, blocksize, 1, 12};
    size_t ws_bytes = 0;
    if ((rc = frb_decode_workspace_size(&p, n_frames, &ws_bytes))) return rc;
    if ((rc = C.d_ws.reserve(ws_bytes))) return rc;
    if ((rc = C.d_audio.reserve(total * 4))) return rc;
    if ((rc = C.d_out.reserve(total * 4))) return rc;
    if ((rc = C.d_misc.reserve(256))) return rc;
    frb_decode_stream st = {0, n_bytes, n_samples, 0, sample_rate, 0};
    uint32_t h_status[8];
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = frb_decode_batch(&p, &st, (const uint8_t *)C.d_in.ptr, n_frames, (int32_t *)C.d_audio.ptr, C.d_ws.ptr, C.d_ws.cap,
                                   (uint32_t *)C.d_misc.ptr, s))) return rc;
        FRB_CUDA(cudaMemcpyAsync(h_status, C.d_misc.ptr, sizeof h_status, cudaMemcpyDeviceToHost, s));
        FRB_CUDA(cudaStreamSynchronize(s));
        if (h_status[4] && attempt == 0) { p.reserved = 32; continue; }   // LPC order > 12: rerun the wide kernel
        break;
    }
    if (h_status[5]) return FRB_ERR_CUDA;                       // a decode thread gave up waiting for its subframe offset
    if (h_status[0] || h_status[2]) return FRB_ERR_BAD_STREAM;
    if (h_status[1]) return FRB_ERR_CRC;
    if (channels == 1) {
        FRB_CUDA(cudaMemcpyAsync(interleaved_out, C.d_audio.ptr, total * 4, cudaMemcpyDeviceToHost, s));
    } else {
        k_interleave<<<grid_for(total, 256 * 8, kNumSMs * 16), 256, 0, s>>>((const int32_t *)C.d_audio.ptr, (int32_t *)C.d_out.ptr, n_samples, channels);
        FRB_LAUNCH_CHECK("k_interleave");
        FRB_CUDA(cudaMemcpyAsync(interleaved_out, C.d_out.ptr, total * 4, cudaMemcpyDeviceToHost, s));
    }
    FRB_CUDA(cudaStreamSynchronize(s));
    return FRB_OK;
}

// ---------------------------------------------------------------- handle API
struct frb_stream_encoder {
    uint32_t channels = 2, bps = 16, sample_rate = 44100, level = 5, blocksize = 0;
    uint64_t total_estimate = 0;
    frb_encoder_write_cb cb = nullptr;
    void *client = nullptr;
    int state = 1;                 // 0 OK (initialised), 1 UNINITIALIZED, 2 error (loosely FLAC__StreamEncoderState)
    std::vector<int32_t> pending;  // interleaved samples not yet forming a full block
    uint64_t frames_out = 0;
    std::vector<uint8_t> outbuf;
    std::vector<uint32_t> fsizes;
};

static const char kVendor[] = "flac-raster-b200 0.1 (sm_100a CUDA FLAC engine)";

extern "C" frb_stream_encoder *frb_stream_encoder_new(void) { return new (std::nothrow) frb_stream_encoder(); }
extern "C" void frb_stream_encoder_delete(frb_stream_encoder *e) { delete e; }
#define FRB_SETTER(name, field, type)                                                       \
    extern "C" int frb_stream_encoder_set_##name(frb_stream_encoder *e, type v) {            \
        if (!e || e->state != 1) return 0;                                                  \
        e->field = v; return 1;                                                             \
    }
FRB_SETTER(channels, channels, uint32_t)
FRB_SETTER(bits_per_sample, bps, uint32_t)
FRB_SETTER(sample_rate, sample_rate, uint32_t)
FRB_SETTER(compression_level, level, uint32_t)
FRB_SETTER(blocksize, blocksize, uint32_t)
FRB_SETTER(total_samples_estimate, total_estimate, uint64_t)
#undef FRB_SETTER
extern "C" int frb_stream_encoder_get_state(const frb_stream_encoder *e) { return e ? e->state : 2; }

extern "C" int frb_stream_encoder_init_stream(frb_stream_encoder *e, frb_encoder_write_cb write_cb, void *client_data) {
    if (!e || !write_cb || e->state != 1) return 1;          // FLAC__STREAM_ENCODER_INIT_STATUS_ENCODER_ERROR-like nonzero
    if (e->blocksize == 0) e->blocksize = 4096;              // libFLAC picks 4096 for LPC presets
    if (e->level > 8) e->level = 8;
    frb_encode_params p = {1, e->channels, e->bps, e->blocksize, e->level, 0};
    if (!frb::enc_params_ok(&p)) return 1;
    e->cb = write_cb; e->client = client_data;
    // "fLaC" + STREAMINFO (unfinalised, like the reference: SURVEY Q7) + VORBIS_COMMENT(vendor only)
    uint8_t hdr[4 + 4 + 34];
    memcpy(hdr, "fLaC", 4);
    uint8_t *si = hdr + 4; memset(si, 0, 38);
    si[3] = 34;
    uint8_t *s = si + 4;
    s[0] = (uint8_t)(e->blocksize >> 8); s[1] = (uint8_t)e->blocksize; s[2] = s[0]; s[3] = s[1];
    s[10] = (uint8_t)(e->sample_rate >> 12); s[11] = (uint8_t)(e->sample_rate >> 4);
    s[12] = (uint8_t)(((e->sample_rate & 15) << 4) | ((e->channels - 1) << 1) | (((e->bps - 1) >> 4) & 1));
    s[13] = (uint8_t)(((e->bps - 1) & 15) << 4);
    if (write_cb(e, hdr, 4, 0, 0, client_data)) { e->state = 2; return 1; }
    if (write_cb(e, hdr + 4, 38, 0, 0, client_data)) { e->state = 2; return 1; }
    const size_t vlen = sizeof(kVendor) - 1;
    std::vector<uint8_t> vc(4 + 4 + vlen + 4, 0);
    vc[0] = 0x84; vc[1] = (uint8_t)((8 + vlen) >> 16); vc[2] = (uint8_t)((8 + vlen) >> 8); vc[3] = (uint8_t)(8 + vlen);
    vc[4] = (uint8_t)vlen; vc[5] = (uint8_t)(vlen >> 8);
    memcpy(vc.data() + 8, kVendor, vlen);
    if (write_cb(e, vc.data(), vc.size(), 0, 0, client_data)) { e->state = 2; return 1; }
    e->state = 0;
    return 0;
}

// All samples are buffered until finish(): the engine encodes a whole stream as one batch (the
// reference makes exactly one process() call followed by finish(), converter.py:153-154).
extern "C" int frb_stream_encoder_process_interleaved(frb_stream_encoder *e, const int32_t *buffer, uint32_t samples) {
    if (!e || e->state != 0 || (!buffer && samples)) return 0;
    e->pending.insert(e->pending.end(), buffer, buffer + (size_t)samples * e->channels);
    return 1;
}

extern "C" int frb_stream_encoder_finish(frb_stream_encoder *e) {
    if (!e) return 0;
    if (e->state != 0) { e->state = 1; return 0; }
    int ok = 1;
    const uint64_t n = e->pending.size() / e->channels;
    if (n) {
        size_t cap = (size_t)n * e->channels * 5 + ((size_t)n / e->blocksize + 2) * 64 + 4096;
        e->outbuf.resize(cap);
        const size_t nf = (size_t)((n + e->blocksize - 1) / e->blocksize);
        e->fsizes.resize(nf);
        size_t ob = 0, got = 0;
        int rc = frb_host_encode(e->pending.data(), n, e->channels, e->bps, e->sample_rate, e->level, e->blocksize, 0,
                                 e->outbuf.data(), cap, &ob, e->fsizes.data(), nf, &got);
        if (rc != FRB_OK) { ok = 0; e->state = 2; }
        else {
            size_t off = 0;
            for (size_t f = 0; f < nf; f++) {
                uint32_t bs = (f + 1 < nf) ? e->blocksize : (uint32_t)(n - (uint64_t)f * e->blocksize);
                if (e->cb(e, e->outbuf.data() + off, e->fsizes[f], bs, (uint32_t)f, e->client)) { ok = 0; break; }
                off += e->fsizes[f];
            }
        }
    }
    e->pending.clear(); e->pending.shrink_to_fit();
    if (e->state == 0) e->state = 1;
    return ok;
}
