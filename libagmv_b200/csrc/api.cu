// C-ABI of libagmv_b200.so (include/agmv_b200.h): context, workspaces and the
// host-side orchestration of the sm_100a kernels. No CPU fallback: every
// compute entry point launches kernels on the context's stream.
#include "../../include/agmv_b200.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>

#include "audio.cuh"
#include "common.cuh"
#include "decode.cuh"
#include "encode.cuh"
#include "lzss.cuh"
#include "lz77.cuh"
#include "radix.cuh"
#include "scan.cuh"

using namespace agmvb;

namespace {

struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct DecStream {
    bool open = false;
    uint32_t w = 0, h = 0, n_frames = 0, next = 0, version = 0;
    int dual = 0, lz77 = 0;
    uint8_t* d_file = nullptr;
    uint64_t file_len = 0;
    std::vector<uint64_t> data_off;
    std::vector<uint32_t> usize, csize;
    std::vector<uint64_t> audio_off;   // per frame: file offset of the 'AGAC' chunk's sample bytes (streams with an audio track)
    std::vector<uint32_t> audio_len;
    uint32_t audio_bits = 16;
    uint32_t* d_pal = nullptr;      // 512
    uint32_t* d_img = nullptr;      // P: pixels after the last decoded frame
    uint32_t* d_ifr = nullptr;      // P: I-frame snapshot
    uint8_t* d_persist = nullptr;   // carried-over expanded bitstream
    uint32_t persist_len = 0;
    uint32_t* d_ring = nullptr;     // optional ring of frames for agmvb_dec_batch without outputs
    uint32_t last_bpos = 0, last_consumed = 0;
    bool raw = false;               // fed chunk by chunk (agmvb_dec_chunk / agmvb_dec_chunks): the vectors hold the chunks of one call
    uint32_t raw_base = 0;          // frame_count of the first chunk of that call
    uint64_t file_cap = 0;
    uint8_t* d_snap = nullptr;      // agmvb_dec_snapshot: pixels, I-frame snapshot, carried bitstream buffer
    bool snap_valid = false;
    size_t at(uint32_t g) const { return raw ? g - raw_base : g; }
};

}  // namespace

struct agmvb_ctx {
    int device = 0;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    char err[512] = {0};
    LaunchCtx lc;

    // ---- encoder state ----
    bool enc_ready = false;
    uint32_t src_w = 0, src_h = 0, cw = 0, ch = 0, mc = 0;
    int opt = 0, quality = 0, compression = 0, dual = 1, light = 0;
    unsigned long long* d_hist = nullptr;
    unsigned long long* d_keys[2] = {nullptr, nullptr};
    uint32_t* d_pal = nullptr;  // 512
    uint32_t h_pal[512];
    bool pal_valid = false;
    uint16_t* d_lut = nullptr;  // 2^24
    bool lut_full = false;      // every colour's entry is in the table (lut_fill_k)
    uint32_t* d_map = nullptr;  // coded pixel -> source pixel (GBA / NDS profiles)
    uint16_t* d_ient = nullptr; // persistent I-frame entries
    size_t ient_px = 0;
    DBuf stage, entries, rec, boff, bs, fs, image, srcpairs, entpairs, scanws, small, seqbuf;
    LzWork lz;
    DBuf lzbuf[64];
    int host_fmt = 0;   // pixel format of HOST frame buffers: 0 = u32 0x00RRGGBB, 1 = packed B,G,R bytes (a BMP's pixel rows)
    DBuf raw24;         // staging for packed 24-bit pixels
    DBuf l77_out, l77_meta, l77_persist, l77_a1, l77_a2, l77_inv, l77_hist, l77_scan, l77_tab;  // LZ77: token words, per-frame arrays, the reference's carried bitstream buffer
    uint64_t image_bytes = 0;
    std::vector<uint32_t> last_usize, last_csize;
    // audio track of the sequence being encoded: the header fields AGMV_WavToAudioTrack sets and the companded bytes
    DBuf at_pcm, at_sample, at_idx, at_lut;
    uint64_t at_size = 0;      // header.audio_size (samples); 0 = no track
    uint32_t at_bits = 16, at_rate = 0, at_channels = 0, at_duration = 0;
    void* h_pinned = nullptr;  // small pinned scratch for async size read-backs
    size_t h_pinned_cap = 0;

    // ---- decoder state ----
    std::vector<DecStream> streams;
    std::vector<DecStream> parked;   // buffers of closed streams, reused by later opens
    DBuf d_frames, d_ebuf, d_bpos, d_consumed, d_stale, d_recs, d_steps, d_out, d_cksum, d_count;
    DBuf d_keeps;   // per (stream, block) of a chunk: some frame's block keeps pixels of the frame before it (index_walk)
    uint32_t* info_bpos = nullptr;      // host arrays a single-stream call wants filled per frame (agmvb_dec_chunks)
    uint32_t* info_consumed = nullptr;
    DBuf d_code, d_segs, d_seglen, d_oexit, d_ow, d_oentry, d_ocum, d_ofinal;
};

#define CK(expr) AGMVB_CUDA_OK(expr)
#define FAIL(code, ...)                                 \
    do {                                                \
        snprintf(ctx->err, sizeof ctx->err, __VA_ARGS__); \
        return code;                                    \
    } while (0)
#define TRY(expr)                 \
    do {                          \
        int _rc = (expr);         \
        if (_rc != OK) return _rc; \
    } while (0)

// Bulk host<->device copies go out in pieces: a copy engine serves its queue in order, so one multi-GB transfer would hold
// up the small parameter copies of every other context (stream) sharing the device for its whole duration.
static cudaError_t copy_pieces(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    const size_t piece = 32ull << 20;
    for (size_t o = 0; o < bytes; o += piece) {
        cudaError_t e = cudaMemcpyAsync(static_cast<uint8_t*>(dst) + o, static_cast<const uint8_t*>(src) + o, std::min(piece, bytes - o), kind, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Transfer gate: the bulk host<->device transfers of the contexts that share a device (a whole sequence up, its decoded frames
// down) go through the link one at a time per direction, first come first served. Several sequences in flight otherwise copy
// at the same time, each at a fraction of the link rate, finish together, then compute together, then download together - they
// stay in step, and link, SMs and the other direction take turns idling (tools/pcie_probe.py: 55.6 GB/s up, 57.3 down, 95-99
// both at once on this pool; four ungated sequences moved 51 GB/s in total). Gated, a sequence uploads at the full rate while
// the others compute and download, and the upload direction - the larger one - stays busy. AGMVB_XFER_GATE=0 disables it.
struct XferGate {
    std::mutex m[2];   // 0: host -> device, 1: device -> host
};
static XferGate& xfer_gate(int device) {
    static XferGate gates[64];
    return gates[device & 63];
}
static bool xfer_gate_on() {
    static const bool on = !(getenv("AGMVB_XFER_GATE") && atoi(getenv("AGMVB_XFER_GATE")) == 0);
    return on;
}
// everything queued on the context's stream so far completes first (so the gate is held for the copy alone), then the copy
// runs to completion inside the gate
struct GatedXfer {
    std::unique_lock<std::mutex> lk;
    GatedXfer(agmvb_ctx* ctx, int dir, cudaStream_t st) {
        if (!xfer_gate_on()) return;
        cudaStreamSynchronize(st);
        lk = std::unique_lock<std::mutex>(xfer_gate(ctx->device).m[dir]);
    }
    cudaError_t finish(cudaStream_t st) {   // call after queueing the copy
        if (!lk.owns_lock()) return cudaSuccess;
        cudaError_t e = cudaStreamSynchronize(st);
        lk.unlock();
        return e;
    }
};

static int ensure(agmvb_ctx* ctx, DBuf& b, size_t bytes) {
    if (b.cap >= bytes && b.p) return OK;
    if (b.p) { CK(cudaStreamSynchronize(ctx->st)); CK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 4 + 256;
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return OK;
}

static int ensure_pinned(agmvb_ctx* ctx, size_t bytes) {
    if (ctx->h_pinned_cap >= bytes) return OK;
    if (ctx->h_pinned) { CK(cudaStreamSynchronize(ctx->st)); CK(cudaFreeHost(ctx->h_pinned)); }
    CK(cudaMallocHost(&ctx->h_pinned, bytes * 2 + 4096));
    ctx->h_pinned_cap = bytes * 2 + 4096;
    return OK;
}

static int check_launch(agmvb_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) FAIL(ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return OK;
}

// ===========================================================================
// context
// ===========================================================================
extern "C" int agmvb_create(agmvb_ctx** out, int device, void* cuda_stream) {
    if (!out) return ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ERR_CUDA;
    agmvb_ctx* ctx = new agmvb_ctx();
    ctx->device = device;
    if (cuda_stream) ctx->st = (cudaStream_t)cuda_stream;
    else {
        if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ERR_CUDA; }
        ctx->own_stream = true;
    }
    ctx->lc.st = ctx->st;
    if (cudaFuncSetAttribute(pal_pick_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536) != cudaSuccess) {
        delete ctx;
        return ERR_CUDA;
    }
    *out = ctx;
    return OK;
}

constexpr uint32_t DEC_RING = 8;   // frames of a stream's output ring (agmvb_dec_batch without output buffers)

static void free_stream(DecStream& s) {
    cudaFree(s.d_file); cudaFree(s.d_pal); cudaFree(s.d_img); cudaFree(s.d_ifr); cudaFree(s.d_persist); cudaFree(s.d_ring); cudaFree(s.d_snap);
    s = DecStream();
}

// Closing a stream parks its device buffers so that the next open of a stream of the same shape does not pay
// cudaMalloc / cudaFree (both synchronise the device).
static void park_stream(agmvb_ctx* ctx, DecStream& s);
static bool unpark_stream(agmvb_ctx* ctx, DecStream& s, uint64_t file_bytes, size_t P);

extern "C" void agmvb_destroy(agmvb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    cudaFree(ctx->d_hist); cudaFree(ctx->d_keys[0]); cudaFree(ctx->d_keys[1]); cudaFree(ctx->d_pal); cudaFree(ctx->d_lut);
    cudaFree(ctx->d_map); cudaFree(ctx->d_ient);
    DBuf* bufs[] = {&ctx->stage, &ctx->entries, &ctx->rec, &ctx->boff, &ctx->bs, &ctx->fs, &ctx->image, &ctx->srcpairs, &ctx->entpairs,
                    &ctx->scanws, &ctx->small, &ctx->seqbuf, &ctx->d_frames, &ctx->d_ebuf, &ctx->d_bpos, &ctx->d_consumed, &ctx->d_stale, &ctx->d_recs,
                    &ctx->d_steps, &ctx->d_out, &ctx->d_cksum, &ctx->d_count, &ctx->d_code, &ctx->d_segs, &ctx->d_seglen, &ctx->d_oexit, &ctx->d_ow,
                    &ctx->d_oentry, &ctx->d_ocum, &ctx->d_ofinal, &ctx->d_keeps};
    for (DBuf* b : bufs) cudaFree(b->p);
    for (DBuf& b : ctx->lzbuf) cudaFree(b.p);
    cudaFree(ctx->at_pcm.p); cudaFree(ctx->at_sample.p); cudaFree(ctx->at_idx.p); cudaFree(ctx->at_lut.p);
    cudaFree(ctx->l77_out.p); cudaFree(ctx->l77_meta.p); cudaFree(ctx->l77_persist.p); cudaFree(ctx->raw24.p);
    cudaFree(ctx->l77_a1.p); cudaFree(ctx->l77_a2.p); cudaFree(ctx->l77_inv.p); cudaFree(ctx->l77_hist.p); cudaFree(ctx->l77_scan.p); cudaFree(ctx->l77_tab.p);
    for (DecStream& s : ctx->streams) if (s.open) free_stream(s);
    for (DecStream& s : ctx->parked) free_stream(s);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->own_stream) cudaStreamDestroy(ctx->st);
    delete ctx;
}

extern "C" const char* agmvb_last_error(const agmvb_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" uint64_t agmvb_kernel_launches(const agmvb_ctx* ctx) { return ctx ? ctx->lc.launches : 0; }
extern "C" int agmvb_sync(agmvb_ctx* ctx) {
    if (!ctx) return ERR_ARG;
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

// ---- host pixel formats ------------------------------------------------------------------------
// The reference's frames come from (and go to) 24-bit BMP files; handing the packed B,G,R rows across PCIe instead of
// 4-byte pixels moves a quarter fewer bytes, and the unpack / pack runs at HBM speed on the device.
__global__ void __launch_bounds__(256) unpack_bgr24_k(const uint8_t* __restrict__ src, uint64_t npx, uint32_t* __restrict__ dst) {
    // four pixels (12 bytes = three aligned words) per thread
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t p0 = g * 4;
    if (p0 >= npx) return;
    if (p0 + 4 <= npx) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(src) + g * 3;
        const uint32_t a = w[0], b = w[1], c = w[2];
        uint4 o;
        o.x = a & 0xFFFFFFu;
        o.y = (a >> 24) | (b & 0xFFFFu) << 8;
        o.z = (b >> 16) | (c & 0xFFu) << 16;
        o.w = c >> 8;
        *reinterpret_cast<uint4*>(dst + p0) = o;   // byte order B,G,R == 0x00RRGGBB little-endian
    } else {
        for (uint64_t p = p0; p < npx; p++) dst[p] = (uint32_t)src[3 * p] | (uint32_t)src[3 * p + 1] << 8 | (uint32_t)src[3 * p + 2] << 16;
    }
}
__global__ void __launch_bounds__(256) pack_bgr24_k(const uint32_t* __restrict__ src, uint64_t npx, uint8_t* __restrict__ dst) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t p0 = g * 4;
    if (p0 >= npx) return;
    if (p0 + 4 <= npx) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + p0);
        uint32_t* w = reinterpret_cast<uint32_t*>(dst) + g * 3;
        w[0] = (v.x & 0xFFFFFFu) | v.y << 24;
        w[1] = ((v.y >> 8) & 0xFFFFu) | v.z << 16;
        w[2] = ((v.z >> 16) & 0xFFu) | v.w << 8;
    } else {
        for (uint64_t p = p0; p < npx; p++) { dst[3 * p] = (uint8_t)src[p]; dst[3 * p + 1] = (uint8_t)(src[p] >> 8); dst[3 * p + 2] = (uint8_t)(src[p] >> 16); }
    }
}

extern "C" int agmvb_set_host_format(agmvb_ctx* ctx, int fmt) {
    if (!ctx || (fmt != 0 && fmt != 1)) return ERR_ARG;
    ctx->host_fmt = fmt;
    return OK;
}

// count host frames (fpx pixels each) starting at frame `first` of the caller's buffer -> d_dst as u32 pixels
static int upload_frames(agmvb_ctx* ctx, const uint32_t* host_frames, uint64_t first, uint64_t count, uint64_t fpx, uint32_t* d_dst) {
    if (ctx->host_fmt == 0) {
        CK(copy_pieces(d_dst, host_frames + first * fpx, count * fpx * 4, cudaMemcpyHostToDevice, ctx->st));
        return OK;
    }
    const uint8_t* src = reinterpret_cast<const uint8_t*>(host_frames) + first * fpx * 3;
    // through a bounded staging buffer so that a long sequence does not need a second full-size copy on the device
    const uint64_t chunk = std::max<uint64_t>(1, (512ull << 20) / (fpx * 3));
    TRY(ensure(ctx, ctx->raw24, std::min(chunk, count) * fpx * 3 + 16));
    for (uint64_t f0 = 0; f0 < count; f0 += chunk) {
        const uint64_t nf = std::min(chunk, count - f0), npx = nf * fpx;
        CK(copy_pieces(ctx->raw24.p, src + f0 * fpx * 3, npx * 3, cudaMemcpyHostToDevice, ctx->st));
        KL(ctx->lc, KC_MISC, (unpack_bgr24_k<<<(unsigned)cdiv(cdiv(npx, 4), 256), 256, 0, ctx->st>>>(ctx->raw24.as<uint8_t>(), npx, d_dst + f0 * fpx)));
    }
    return OK;
}

// ===========================================================================
// encoder
// ===========================================================================
extern "C" int agmvb_enc_begin(agmvb_ctx* ctx, uint32_t src_w, uint32_t src_h, int opt, int quality, int compression) {
    if (!ctx) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (opt < OPT_I || opt > OPT_NDS || quality < Q_HIGH || quality > Q_LOW) FAIL(ERR_ARG, "bad opt/quality");
    if (compression != COMP_LZSS && compression != COMP_LZ77) FAIL(ERR_ARG, "bad compression %d", compression);
    ctx->src_w = src_w; ctx->src_h = src_h; ctx->opt = opt; ctx->quality = quality; ctx->compression = compression;
    ctx->dual = opt_is_dual(opt); ctx->light = opt_is_light(opt);
    ctx->cw = src_w; ctx->ch = src_h;
    if (opt == OPT_GBA_I || opt == OPT_GBA_II || opt == OPT_GBA_III) { ctx->cw = 120; ctx->ch = 80; }  // include/agmv_encode.h:21-24
    if (opt == OPT_NDS) { ctx->cw = 128; ctx->ch = 96; }
    if (ctx->cw == 0 || ctx->ch == 0 || (ctx->cw & 3) || (ctx->ch & 3) || (src_w & 3))
        FAIL(ERR_UNSUPPORTED, "width and height must be multiples of 4 (got %ux%u, coded %ux%u)", src_w, src_h, ctx->cw, ctx->ch);
    ctx->mc = max_clr(quality);
    const size_t P = (size_t)ctx->cw * ctx->ch;
    if (!ctx->d_hist) CK(cudaMalloc(&ctx->d_hist, sizeof(unsigned long long) * (524287 + 1)));
    if (!ctx->d_keys[0]) { CK(cudaMalloc(&ctx->d_keys[0], 8 * 524288)); CK(cudaMalloc(&ctx->d_keys[1], 8 * 524288)); }
    if (!ctx->d_pal) CK(cudaMalloc(&ctx->d_pal, 512 * 4));
    if (!ctx->d_lut) CK(cudaMalloc(&ctx->d_lut, sizeof(uint16_t) << 24));
    CK(cudaMemsetAsync(ctx->d_hist, 0, sizeof(unsigned long long) * (524287 + 1), ctx->st));
    CK(cudaMemsetAsync(ctx->d_lut, 0xFF, sizeof(uint16_t) << 24, ctx->st));
    ctx->lut_full = false;
    if (ctx->d_ient && ctx->ient_px != P) { CK(cudaStreamSynchronize(ctx->st)); CK(cudaFree(ctx->d_ient)); ctx->d_ient = nullptr; }
    if (!ctx->d_ient) { CK(cudaMalloc(&ctx->d_ient, P * 2)); ctx->ient_px = P; }
    CK(cudaMemsetAsync(ctx->d_ient, 0, P * 2, ctx->st));
    if (ctx->d_map) { CK(cudaStreamSynchronize(ctx->st)); CK(cudaFree(ctx->d_map)); ctx->d_map = nullptr; }
    if (ctx->cw != src_w || ctx->ch != src_h) {
        // AGIDL_ScaleImgDataNearest (extern/agidl/src/agidl_imgp_scale.c:262-293), called with
        // sx = (f32)W/w + 0.001f (src/agmv_encode.c:2707-2721). The float32 index math runs here,
        // on the host, once per sequence; the kernel just gathers through the table.
        unsigned long worg = src_w, horg = src_h;
        float sx = ((float)ctx->cw / worg) + 0.001f, sy = ((float)ctx->ch / horg) + 0.001f;
        unsigned long neww = (unsigned long)(worg * sx), newh = (unsigned long)(horg * sy);
        if (neww != ctx->cw || newh != ctx->ch) FAIL(ERR_UNSUPPORTED, "source %ux%u does not scale to %ux%u", src_w, src_h, ctx->cw, ctx->ch);
        float xscale = (float)(worg - 1) / neww, yscale = (float)(horg - 1) / newh;
        std::vector<uint32_t> map(P);
        for (unsigned long y = 0; y < newh; y++)
            for (unsigned long x = 0; x < neww; x++) {
                unsigned long x2 = (unsigned long)(x * xscale), y2 = (unsigned long)(y * yscale);
                map[y * neww + x] = (uint32_t)(y2 * worg + x2);
            }
        CK(cudaMalloc(&ctx->d_map, P * 4));
        CK(cudaMemcpyAsync(ctx->d_map, map.data(), P * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    }
    if (compression == COMP_LZ77) {
        // CreateAGMV's bitstream buffer (src/agmv_utils.c:338): fresh zero pages, never cleared between frames
        const size_t pb = (P / 16) * 33 + 64;
        TRY(ensure(ctx, ctx->l77_persist, pb));
        CK(cudaMemsetAsync(ctx->l77_persist.p, 0, pb, ctx->st));
    }
    ctx->pal_valid = false;
    ctx->image_bytes = 0;
    ctx->enc_ready = true;
    return OK;
}

static int hist_launch(agmvb_ctx* ctx, const uint32_t* d_px, uint64_t npx) {
    if (npx == 0) return OK;
    if (((uintptr_t)d_px & 15) == 0) {
        uint64_t groups = npx / 4;
        if (groups) {
            int grid = (int)std::min<uint64_t>((groups + 255) / 256, 148 * 16);
            KL(ctx->lc, KC_HIST, (hist_vec4_k<<<grid, 256, 0, ctx->st>>>(reinterpret_cast<const uint4*>(d_px), groups, ctx->quality, ctx->mc, ctx->d_hist)));
        }
        uint64_t tail = npx - groups * 4;
        if (tail) {
            KL(ctx->lc, KC_HIST, (hist_scalar_k<<<1, 256, 0, ctx->st>>>(d_px + groups * 4, tail, ctx->quality, ctx->mc, ctx->d_hist)));
        }
    } else {
        int grid = (int)std::min<uint64_t>((npx + 255) / 256, 148 * 16);
        KL(ctx->lc, KC_HIST, (hist_scalar_k<<<grid, 256, 0, ctx->st>>>(d_px, npx, ctx->quality, ctx->mc, ctx->d_hist)));
    }
    return check_launch(ctx, "histogram");
}

extern "C" int agmvb_enc_histogram(agmvb_ctx* ctx, const uint32_t* frames, uint64_t n_frames, int on_device) {
    if (!ctx || !ctx->enc_ready) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint64_t fpx = (uint64_t)ctx->src_w * ctx->src_h;
    if (on_device) return hist_launch(ctx, frames, fpx * n_frames);
    // host frames: stream them through a staging buffer
    const uint64_t chunk_frames = std::max<uint64_t>(1, (256ull << 20) / (fpx * 4));
    TRY(ensure(ctx, ctx->stage, std::min(chunk_frames, n_frames) * fpx * 4));
    for (uint64_t f0 = 0; f0 < n_frames; f0 += chunk_frames) {
        uint64_t nf = std::min(chunk_frames, n_frames - f0);
        TRY(upload_frames(ctx, frames, f0, nf, fpx, ctx->stage.as<uint32_t>()));
        TRY(hist_launch(ctx, ctx->stage.as<uint32_t>(), nf * fpx));
        CK(cudaStreamSynchronize(ctx->st));  // the staging buffer is reused
    }
    return OK;
}

extern "C" int agmvb_enc_histogram_ptr(agmvb_ctx* ctx, uint64_t** dev_bins, uint32_t* n_bins) {
    if (!ctx || !ctx->enc_ready) return ERR_ARG;
    if (dev_bins) *dev_bins = reinterpret_cast<uint64_t*>(ctx->d_hist);
    if (n_bins) *n_bins = ctx->mc;
    return OK;
}

extern "C" int agmvb_enc_build_palette(agmvb_ctx* ctx) {
    if (!ctx || !ctx->enc_ready) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint32_t mc = ctx->mc;
    KL(ctx->lc, KC_PALETTE, (pal_keys_k<<<cdiv(mc, 256), 256, 0, ctx->st>>>(ctx->d_hist, mc, ctx->d_keys[0])));
    const uint32_t nt = cdiv(mc, RX_TILE);
    TRY(ensure(ctx, ctx->small, (size_t)256 * nt * 4));
    TRY(ensure(ctx, ctx->scanws, ((size_t)cdiv((size_t)256 * nt, SCAN_TILE) + 2) * 4));
    int cur = 0;
    for (uint32_t shift = 0; shift < 64; shift += 8) {  // stable LSD passes == stable sort by (count, index)
        radix_pass(KeyDigit{ctx->d_keys[cur], shift}, KeyMove{ctx->d_keys[cur], ctx->d_keys[cur ^ 1]}, mc, ctx->small.as<uint32_t>(),
                   ctx->scanws.as<uint32_t>(), ctx->lc);
        cur ^= 1;
    }
    KL(ctx->lc, KC_PALETTE, (pal_pick_k<<<1, 32, (mc + 1) / 8, ctx->st>>>(ctx->d_keys[cur], mc, ctx->quality, ctx->dual, ctx->d_pal)));
    TRY(check_launch(ctx, "palette"));
    CK(cudaMemcpyAsync(ctx->h_pal, ctx->d_pal, 512 * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->pal_valid = true;
    return OK;
}

extern "C" int agmvb_enc_get_palette(agmvb_ctx* ctx, uint32_t pal0[256], uint32_t pal1[256]) {
    if (!ctx || !ctx->pal_valid) return ERR_ARG;
    if (pal0) memcpy(pal0, ctx->h_pal, 1024);
    if (pal1) memcpy(pal1, ctx->h_pal + 256, 1024);
    return OK;
}

extern "C" int agmvb_enc_set_palette(agmvb_ctx* ctx, const uint32_t pal0[256], const uint32_t pal1[256]) {
    if (!ctx || !ctx->enc_ready || !pal0) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    bool same = ctx->pal_valid && !memcmp(ctx->h_pal, pal0, 1024) && (!pal1 || !memcmp(ctx->h_pal + 256, pal1, 1024));
    if (same) return OK;
    memcpy(ctx->h_pal, pal0, 1024);
    if (pal1) memcpy(ctx->h_pal + 256, pal1, 1024); else memset(ctx->h_pal + 256, 0, 1024);
    for (int i = 0; i < 512; i++) ctx->h_pal[i] &= 0xFFFFFFu;
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaMemcpyAsync(ctx->d_pal, ctx->h_pal, 512 * 4, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(ctx->d_lut, 0xFF, sizeof(uint16_t) << 24, ctx->st));  // the memo table belongs to a palette
    ctx->lut_full = false;
    CK(cudaStreamSynchronize(ctx->st));
    ctx->pal_valid = true;
    return OK;
}

extern "C" int agmvb_enc_set_iframe_entries(agmvb_ctx* ctx, const uint16_t* entries) {
    if (!ctx || !ctx->enc_ready || !entries) return ERR_ARG;
    CK(cudaMemcpyAsync(ctx->d_ient, entries, (size_t)ctx->cw * ctx->ch * 2, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}
extern "C" int agmvb_enc_get_iframe_entries(agmvb_ctx* ctx, uint16_t* entries) {
    if (!ctx || !ctx->enc_ready || !entries) return ERR_ARG;
    CK(cudaMemcpyAsync(entries, ctx->d_ient, (size_t)ctx->cw * ctx->ch * 2, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

// LZSS workspace for up to n positions and F frames
static int ensure_lz(agmvb_ctx* ctx, uint32_t n, uint32_t F) {
    LzWork& w = ctx->lz;
    const uint32_t cap_before = w.cap_n;
    if (n + 64 > w.cap_n || !w.bestlen) {
        uint32_t cap = std::max<uint32_t>(n + n / 4 + 4096, 1u << 20);
        int k = 0;
        auto grab = [&](size_t bytes, void** dst) -> int {
            TRY(ensure(ctx, ctx->lzbuf[k], bytes));
            *dst = ctx->lzbuf[k].p;
            k++;
            return OK;
        };
        // occurrence-chain match finder (lzchain.cuh): 19 B per position
        TRY(grab(cap, (void**)&w.bestlen));
        TRY(grab((size_t)cap * 4, (void**)&w.lw[0]));
        TRY(grab((size_t)cap * 4, (void**)&w.lw[1]));
        TRY(grab((size_t)cap * 2, (void**)&w.rsd));
        TRY(grab(64 * sizeof(uint32_t), (void**)&w.counters));
        TRY(grab(((size_t)cap / LZC_WCHUNK + 2) * 4, (void**)&w.cframe));
        int sms = 148, per3 = 5, perl = 6;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per3, lzc_link3_k<LZC_ROUNDS>, LZC_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perl, lzc_level_k<LZC_ROUNDS>, LZC_THREADS, 0);
        w.link3_blocks = (uint32_t)(sms * std::max(1, per3));
        w.level_blocks = (uint32_t)(sms * std::max(1, perl));
        w.cap_n = cap;
    }
    size_t words = (((size_t)n * 9) >> 5) + 3 * (size_t)F + 16;
    const bool grew = w.cap_n != cap_before;
    if (words > w.out_words_cap || F + 2 > w.cap_frames || !w.out_words || grew) {
        uint32_t capF = std::max<uint32_t>(F + F / 4 + 16, 64);
        size_t capW = std::max(words, (((size_t)w.cap_n * 9) >> 5) + 3 * (size_t)capF + 16);
        TRY(ensure(ctx, ctx->lzbuf[52], capW * 4)); w.out_words = ctx->lzbuf[52].as<uint32_t>(); w.out_words_cap = capW;
        TRY(ensure(ctx, ctx->lzbuf[53], (size_t)(capF + 2) * 4)); w.wbase = ctx->lzbuf[53].as<uint32_t>();
        TRY(ensure(ctx, ctx->lzbuf[54], (size_t)(capF + 2) * 4)); w.outbits = ctx->lzbuf[54].as<uint32_t>();
        TRY(ensure(ctx, ctx->lzbuf[55], (size_t)(capF + 2) * 4)); w.csize = ctx->lzbuf[55].as<uint32_t>();
        TRY(ensure(ctx, ctx->lzbuf[56], (size_t)(capF + 2) * 4)); w.chunk_off = ctx->lzbuf[56].as<uint32_t>();
        TRY(ensure(ctx, ctx->lzbuf[57], (size_t)(capF + 2) * sizeof(OrbitSeg))); w.segs = ctx->lzbuf[57].as<OrbitSeg>();
        TRY(ensure(ctx, ctx->lzbuf[58], (size_t)(capF + 2) * 4)); w.seg_len = ctx->lzbuf[58].as<uint32_t>();
        TRY(ensure(ctx, ctx->lzbuf[59], (size_t)(capF + 2) * 8)); w.orb.final_pos = ctx->lzbuf[59].as<uint32_t>();
        w.orb.final_cum = w.orb.final_pos + (capF + 2);
        // parse tables: every frame adds at most one partial tile
        const size_t ptile = (size_t)w.cap_n / ORB_TILE + capF + 8;
        TRY(ensure(ctx, ctx->lzbuf[49], ptile * ORB_SP)); w.orb.exit_tab = ctx->lzbuf[49].as<uint8_t>();
        TRY(ensure(ctx, ctx->lzbuf[50], ptile * ORB_SP * 2)); w.orb.w_tab = ctx->lzbuf[50].as<uint16_t>();
        TRY(ensure(ctx, ctx->lzbuf[51], ptile)); w.orb.entry_tab = ctx->lzbuf[51].as<uint8_t>();
        TRY(ensure(ctx, ctx->lzbuf[47], ptile * 4)); w.orb.cumbase = ctx->lzbuf[47].as<uint32_t>();
        TRY(ensure(ctx, ctx->lzbuf[46], ptile * LZ15_SLOTS * 4)); w.list15 = ctx->lzbuf[46].as<uint32_t>();
        TRY(ensure(ctx, ctx->lzbuf[45], ptile)); w.cnt15 = ctx->lzbuf[45].as<uint8_t>();
        w.cap_frames = capF;
    }
    return OK;
}

// Compress F bitstreams that sit back to back in `d_bs` (n bytes, frame starts
// h_fs[0..F], h_fs[0] == 0) and append their chunk images to ctx->image.
static int lz_group(agmvb_ctx* ctx, const uint8_t* d_bs, const uint32_t* h_fs, uint32_t F, uint32_t first_fc) {
    const uint32_t n = h_fs[F];
    TRY(ensure_lz(ctx, n, F));
    TRY(ensure_pinned(ctx, (size_t)(F + 2) * 8));
    uint32_t* hp = reinterpret_cast<uint32_t*>(ctx->h_pinned);
    memcpy(hp, h_fs, (size_t)(F + 1) * 4);
    TRY(ensure(ctx, ctx->fs, (size_t)(F + 2) * 4));
    CK(cudaMemcpyAsync(ctx->fs.p, hp, (size_t)(F + 1) * 4, cudaMemcpyHostToDevice, ctx->st));
    size_t need = ctx->image_bytes + (size_t)(24 + ctx->lz.stub_bytes) * F + (((size_t)n * 9) >> 3) + 64;
    if (need > ctx->image.cap) {  // grow, keeping what is already there
        DBuf nb;
        TRY(ensure(ctx, nb, need + need / 2));
        if (ctx->image_bytes) CK(cudaMemcpyAsync(nb.p, ctx->image.p, ctx->image_bytes, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (ctx->image.p) CK(cudaFree(ctx->image.p));
        ctx->image = nb;
    }
    // parse segments: one per frame
    std::vector<OrbitSeg> segs(F);
    std::vector<uint32_t> slen(F);
    uint32_t ntile = 0, max_usize = 1;
    for (uint32_t f = 0; f < F; f++) {
        slen[f] = h_fs[f + 1] - h_fs[f];
        max_usize = std::max(max_usize, slen[f]);
        segs[f].off = h_fs[f]; segs[f].cap_len = slen[f]; segs[f].tile_base = ntile;
        ntile += orbit_tiles(slen[f]);
    }
    CK(cudaMemcpyAsync(ctx->lz.segs, segs.data(), F * sizeof(OrbitSeg), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->lz.seg_len, slen.data(), F * 4, cudaMemcpyHostToDevice, ctx->st));
    std::vector<LzcItem> items;   // (frame, range) work items of the serial hash-link kernel
    static const int hb_env = getenv("AGMVB_LZ_HB") ? std::min(13, std::max(8, atoi(getenv("AGMVB_LZ_HB")))) : 0;
    if (hb_env) ctx->lz.hash_bits = hb_env;
    lzc_build_items(h_fs, F, items, ctx->lz.hash_bits);
    TRY(ensure(ctx, ctx->lzbuf[60], (items.size() + 1) * sizeof(LzcItem)));
    ctx->lz.items = ctx->lzbuf[60].as<LzcItem>();
    ctx->lz.n_items = (uint32_t)items.size();
    if (!items.empty()) CK(cudaMemcpyAsync(ctx->lz.items, items.data(), items.size() * sizeof(LzcItem), cudaMemcpyHostToDevice, ctx->st));
    lzss_encode_batch(ctx->lz, d_bs, ctx->fs.as<uint32_t>(), F, n, ntile, max_usize, first_fc, ctx->image.as<uint8_t>() + ctx->image_bytes, ctx->lc);
    CK(cudaStreamSynchronize(ctx->st));  // segs / slen are stack vectors
    TRY(check_launch(ctx, "lzss"));
    uint32_t* hcs = hp + (F + 2);
    CK(cudaMemcpyAsync(hcs, ctx->lz.csize, (size_t)F * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    for (uint32_t f = 0; f < F; f++) {
        ctx->last_usize.push_back(h_fs[f + 1] - h_fs[f]);
        ctx->last_csize.push_back(hcs[f]);
        ctx->image_bytes += 24ull + ctx->lz.stub_bytes + hcs[f];
    }
    return OK;
}

// LZ77 flavour of lz_group (AGMV_LZ77, src/agmv_encode.c:179-238): tokens by lz77_encode_k, chunk framing shared with LZSS
static int lz77_group(agmvb_ctx* ctx, const uint8_t* d_bs, const uint32_t* h_fs, uint32_t F, uint32_t first_fc) {
    const uint32_t n = h_fs[F];
    // per-frame device arrays: fs, wbase, stale_at, total_bits, outbits, csize, chunk_off (F + 2 words each)
    const size_t stride = (size_t)F + 2;
    TRY(ensure(ctx, ctx->l77_meta, stride * 7 * 4));
    TRY(ensure(ctx, ctx->l77_out, ((size_t)n + F + 16) * 4));
    TRY(ensure_pinned(ctx, stride * 4 * 4));
    uint32_t* d_fs = ctx->l77_meta.as<uint32_t>();
    uint32_t *d_wbase = d_fs + stride, *d_stale = d_wbase + stride, *d_bits = d_stale + stride, *d_outbits = d_bits + stride,
             *d_csize = d_outbits + stride, *d_choff = d_csize + stride;
    uint32_t* hp = reinterpret_cast<uint32_t*>(ctx->h_pinned);
    uint32_t *h_w = hp + stride, *h_st = h_w + stride;
    memcpy(hp, h_fs, (size_t)(F + 1) * 4);
    // data[pos] of frame f: written by the latest earlier frame that was longer; frames of earlier groups live in l77_persist
    std::vector<uint32_t> stack;  // frames with strictly decreasing usize towards the top
    for (uint32_t f = 0; f < F; f++) {
        const uint32_t us = h_fs[f + 1] - h_fs[f];
        h_w[f] = h_fs[f] + f;
        while (!stack.empty() && h_fs[stack.back() + 1] - h_fs[stack.back()] <= us) stack.pop_back();
        h_st[f] = stack.empty() ? 0xFFFFFFFFu : h_fs[stack.back()] + us;
        stack.push_back(f);
    }
    CK(cudaMemcpyAsync(d_fs, hp, stride * 3 * 4, cudaMemcpyHostToDevice, ctx->st));
    size_t need = ctx->image_bytes + (size_t)(24 + ctx->lz.stub_bytes) * F + (size_t)n * 4 + 64;
    if (need > ctx->image.cap) {
        DBuf nb;
        TRY(ensure(ctx, nb, need + need / 2));
        if (ctx->image_bytes) CK(cudaMemcpyAsync(nb.p, ctx->image.p, ctx->image_bytes, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (ctx->image.p) CK(cudaFree(ctx->image.p));
        ctx->image = nb;
    }
    CK(cudaMemsetAsync(ctx->l77_out.p, 0, ((size_t)n + F + 16) * 4, ctx->st));
    if (n == 0) {   // nothing but empty frames: no candidate lists to build, the plain per-frame kernel writes the empty payloads
        KL(ctx->lc, KC_LZ77, (lz77_encode_k<<<F, L77_THREADS, 0, ctx->st>>>(d_bs, d_fs, d_stale, ctx->l77_persist.as<uint8_t>(), d_wbase,
                                                                          ctx->l77_out.as<uint32_t>(), d_bits)));
    } else {
        // candidate lists: positions sorted by (byte 1, byte 0, position), slice tables per 2-byte and per 1-byte prefix
        const uint32_t nt = cdiv(n, RX_TILE);
        TRY(ensure(ctx, ctx->l77_a1, (size_t)n * 4 + 16));
        TRY(ensure(ctx, ctx->l77_a2, (size_t)n * 4 + 16));
        TRY(ensure(ctx, ctx->l77_inv, (size_t)n * 4 + 16));
        TRY(ensure(ctx, ctx->l77_hist, (size_t)256 * nt * 4));
        TRY(ensure(ctx, ctx->l77_scan, ((size_t)cdiv((size_t)256 * nt, SCAN_TILE) + 4) * 4));
        TRY(ensure(ctx, ctx->l77_tab, (size_t)(260 + 2 * 65536) * 4));
        uint32_t *a1 = ctx->l77_a1.as<uint32_t>(), *a2 = ctx->l77_a2.as<uint32_t>(), *th = ctx->l77_hist.as<uint32_t>();
        uint32_t *b1 = ctx->l77_tab.as<uint32_t>(), *b2s = b1 + 260, *b2e = b2s + 65536;
        CK(cudaMemsetAsync(b1, 0, (size_t)(260 + 2 * 65536) * 4, ctx->st));
        radix_pass(L77Byte0{d_bs}, L77Move0{a1}, n, th, ctx->l77_scan.as<uint32_t>(), ctx->lc);
        KL(ctx->lc, KC_LZ77, (l77_b1_k<<<1, 256, 0, ctx->st>>>(th, nt, n, b1)));
        radix_pass(L77Byte1{d_bs, a1}, L77Move1{a1, a2}, n, th, ctx->l77_scan.as<uint32_t>(), ctx->lc);
        KL(ctx->lc, KC_LZ77, (l77_buckets_k<<<cdiv(n, 256u), 256, 0, ctx->st>>>(d_bs, a2, n, b2s, b2e, ctx->l77_inv.as<uint32_t>())));
        // frames in flight per launch (measured: the more the better - 1497 frames in one launch 270 ms, in waves of 592 586 ms)
        static const uint32_t wave = getenv("AGMVB_LZ77_WAVE") ? (uint32_t)std::max(1, atoi(getenv("AGMVB_LZ77_WAVE"))) : (1u << 20);
        for (uint32_t f0 = 0; f0 < F; f0 += wave) {
            const uint32_t nf = std::min(wave, F - f0);
            KL(ctx->lc, KC_LZ77, (lz77_bucket_encode_k<<<nf, T77, 0, ctx->st>>>(d_bs, d_fs + f0, a1, a2, b1, b2s, b2e, ctx->l77_inv.as<uint32_t>(), d_stale + f0,
                                                                            ctx->l77_persist.as<uint8_t>(), d_wbase + f0, ctx->l77_out.as<uint32_t>(), d_bits + f0)));
        }
    }
    KL(ctx->lc, KC_LZ_CHUNK, (lz_finalize_k<<<1, 1024, 0, ctx->st>>>(F, ctx->lz.stub_bytes, d_bits, d_outbits, d_csize, d_choff)));
    dim3 grid(32, F);
    KL(ctx->lc, KC_LZ_CHUNK, (lz_write_chunks_k<<<grid, 256, 0, ctx->st>>>(d_fs, d_csize, d_choff, d_wbase, ctx->l77_out.as<uint32_t>(), first_fc,
                                                                          ctx->lz.stub_bytes, ctx->image.as<uint8_t>() + ctx->image_bytes,
                                                                          ctx->lz.audio_chunk, ctx->lz.audio, ctx->lz.audio_size)));
    // carry the buffer forward: index j now holds the byte of the latest frame longer than j (the stack, bottom = longest)
    {
        uint32_t lo = 0;
        for (size_t k = stack.size(); k-- > 0;) {
            const uint32_t f = stack[k], us = h_fs[f + 1] - h_fs[f];
            if (us > lo) CK(cudaMemcpyAsync(ctx->l77_persist.as<uint8_t>() + lo, d_bs + h_fs[f] + lo, us - lo, cudaMemcpyDeviceToDevice, ctx->st));
            lo = std::max(lo, us);
        }
    }
    uint32_t* hcs = hp + stride * 3;
    CK(cudaMemcpyAsync(hcs, d_csize, (size_t)F * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    TRY(check_launch(ctx, "lz77"));
    for (uint32_t f = 0; f < F; f++) {
        ctx->last_usize.push_back(h_fs[f + 1] - h_fs[f]);
        ctx->last_csize.push_back(hcs[f]);
        ctx->image_bytes += 24ull + ctx->lz.stub_bytes + hcs[f];
    }
    return OK;
}

static const uint32_t LZ_GROUP_TARGET = 1000u << 20;       // positions per LZSS batch: 19 B of workspace each (<= 19 GB)

// classify + assemble n frames whose entries are on the device; runs LZSS group by group
static int assemble_and_compress(agmvb_ctx* ctx, const EntPair* h_pairs, uint32_t F, uint32_t first_fc) {
    const uint32_t W = ctx->cw, H = ctx->ch, B = (W >> 2) * (H >> 2);
    const size_t nb = (size_t)F * B;
    TRY(ensure(ctx, ctx->entpairs, F * sizeof(EntPair)));
    TRY(ensure(ctx, ctx->rec, nb));
    TRY(ensure(ctx, ctx->boff, nb * 4));
    TRY(ensure(ctx, ctx->scanws, ((size_t)cdiv(nb, SCAN_TILE) + 2) * 4));
    TRY(ensure(ctx, ctx->bs, nb * 33 + 256));
    TRY(ensure(ctx, ctx->fs, (size_t)(F + 2) * 4));
    TRY(ensure_pinned(ctx, (size_t)(F + 2) * 8));
    CK(cudaMemcpyAsync(ctx->entpairs.p, h_pairs, F * sizeof(EntPair), cudaMemcpyHostToDevice, ctx->st));
    dim3 grid(cdiv(B, 256), F);
    KL(ctx->lc, KC_CLASSIFY, (classify_k<<<grid, 256, 0, ctx->st>>>(ctx->entpairs.as<EntPair>(), W, H, ctx->d_pal, ctx->dual, ctx->rec.as<uint8_t>())));
    device_scan<SumOp, true>(RecLen{ctx->rec.as<uint8_t>()}, StoreU32{ctx->boff.as<uint32_t>()}, (uint32_t)nb, ctx->scanws.as<uint32_t>(),
                             ctx->lc, KC_BLOCKSCAN);
    KL(ctx->lc, KC_BLOCKSCAN, (frame_starts_k<<<cdiv(F + 1, 256), 256, 0, ctx->st>>>(ctx->boff.as<uint32_t>(), B, F, ctx->scanws.as<uint32_t>() + cdiv(nb, SCAN_TILE),
                                                          ctx->fs.as<uint32_t>())));
    KL(ctx->lc, KC_EMIT, (emit_k<<<grid, 256, 0, ctx->st>>>(ctx->entpairs.as<EntPair>(), W, H, ctx->dual, ctx->rec.as<uint8_t>(), ctx->boff.as<uint32_t>(),
                                      ctx->bs.as<uint8_t>())));
    TRY(check_launch(ctx, "assemble"));
    std::vector<uint32_t> fs(F + 1);
    CK(cudaMemcpyAsync(ctx->h_pinned, ctx->fs.p, (size_t)(F + 1) * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    memcpy(fs.data(), ctx->h_pinned, (size_t)(F + 1) * 4);
    std::vector<uint32_t> rebased;
    for (uint32_t g0 = 0; g0 < F;) {
        uint32_t g1 = g0 + 1;
        // LZ77 keeps 4 B of workspace per position and wants every frame of the batch in flight at once (one CTA per frame)
        static const uint32_t lz_target_env = getenv("AGMVB_LZ_TARGET") ? (uint32_t)std::max(1, atoi(getenv("AGMVB_LZ_TARGET"))) << 20 : 0u;
        const uint32_t target = ctx->compression == COMP_LZ77 ? (1u << 30) : (lz_target_env ? lz_target_env : LZ_GROUP_TARGET);
        while (g1 < F && fs[g1 + 1] - fs[g0] <= target) g1++;
        rebased.resize(g1 - g0 + 1);
        for (uint32_t k = 0; k <= g1 - g0; k++) rebased[k] = fs[g0 + k] - fs[g0];
        if (ctx->compression == COMP_LZ77) TRY(lz77_group(ctx, ctx->bs.as<uint8_t>() + fs[g0], rebased.data(), g1 - g0, first_fc + g0));
        else TRY(lz_group(ctx, ctx->bs.as<uint8_t>() + fs[g0], rebased.data(), g1 - g0, first_fc + g0));
        g0 = g1;
    }
    return OK;
}

extern "C" int agmvb_enc_frames(agmvb_ctx* ctx, const uint32_t* frames, uint64_t n_frames_in_buffer, int on_device,
                                const int32_t* src_a, const int32_t* src_b, uint32_t n_enc, uint32_t first_frame_count,
                                uint64_t* image_bytes) {
    if (!ctx || !ctx->enc_ready || !frames || !src_a || !src_b) return ERR_ARG;
    if (!ctx->pal_valid) FAIL(ERR_ARG, "no palette: call agmvb_enc_build_palette or agmvb_enc_set_palette first");
    CK(cudaSetDevice(ctx->device));
    const uint32_t W = ctx->cw, H = ctx->ch;
    const size_t P = (size_t)W * H, SP = (size_t)ctx->src_w * ctx->src_h;
    ctx->image_bytes = 0;
    ctx->last_usize.clear();
    ctx->last_csize.clear();
    for (uint32_t k = 0; k < n_enc; k++)
        if (src_a[k] < 0 || (uint64_t)src_a[k] >= n_frames_in_buffer || (src_b[k] >= 0 && (uint64_t)src_b[k] >= n_frames_in_buffer))
            FAIL(ERR_ARG, "frame index out of range at encoded frame %u", k);
    // frames per quantise batch. LZ77 parses one frame per CTA, serially: it wants every frame of the call in flight at once
    // (entries 2 B/px + bitstream <= 2.1 B/px per frame: ~17 GB for 2000 1080p frames)
    // The assembled bitstream of a batch is addressed with 32-bit offsets (block offsets, frame starts): a block record is at
    // most 33 bytes, so F * B * 33 must stay below 2^32 whatever the content.
    const size_t B33 = (P / 16) * 33 + 1;
    const size_t qb_off = std::max<size_t>(1, (size_t)0xFFFFFFFFull / B33);
    if (qb_off < 4 && n_enc > qb_off) FAIL(ERR_UNSUPPORTED, "frame of %u x %u pixels: bitstream offsets do not fit 32 bits", W, H);
    static const size_t qb_env = getenv("AGMVB_QB") ? (size_t)std::max(1, atoi(getenv("AGMVB_QB"))) : 0;
    size_t qb = ctx->compression == COMP_LZ77 ? std::max<size_t>(4, std::min<size_t>(4096, (24ull << 30) / (P * 5)))
                                              : std::max<size_t>(4, std::min<size_t>(1024, (4096ull << 20) / (P * 2)));
    if (qb_env) qb = qb_env;
    const uint32_t QB = (uint32_t)std::min(qb, qb_off);
    // enough pixels ahead to pay for the whole colour table at once (about 1 ms, encode.cuh lut_fill_k): no quantiser launch of
    // this palette meets an empty entry afterwards. AGMVB_LUT_FILL_PX overrides the threshold (pixels of the call; 0 = always).
    static const uint64_t fill_from = getenv("AGMVB_LUT_FILL_PX") ? strtoull(getenv("AGMVB_LUT_FILL_PX"), nullptr, 10) : (1ull << 27);
    if (!ctx->lut_full && (uint64_t)n_enc * P >= fill_from) {
        KL(ctx->lc, KC_QUANT, (lut_fill_k<<<(1u << 24) / 1024u, 256, 0, ctx->st>>>(ctx->d_pal, ctx->dual ? 512 : 256, reinterpret_cast<uint2*>(ctx->d_lut))));
        TRY(check_launch(ctx, "lut_fill"));
        ctx->lut_full = true;
    }
    std::vector<SrcPair> sp;
    std::vector<EntPair> ep;
    for (uint32_t q0 = 0; q0 < n_enc; q0 += QB) {
        const uint32_t F = std::min(QB, n_enc - q0);
        sp.resize(F);
        ep.resize(F);
        if (on_device) {
            for (uint32_t k = 0; k < F; k++) {
                sp[k].a = frames + (size_t)src_a[q0 + k] * SP;
                sp[k].b = src_b[q0 + k] >= 0 ? frames + (size_t)src_b[q0 + k] * SP : nullptr;
            }
        } else {
            // upload the source frames this batch touches (each at most once)
            std::vector<int32_t> uniq;
            for (uint32_t k = 0; k < F; k++) { uniq.push_back(src_a[q0 + k]); if (src_b[q0 + k] >= 0) uniq.push_back(src_b[q0 + k]); }
            std::sort(uniq.begin(), uniq.end());
            uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
            TRY(ensure(ctx, ctx->stage, uniq.size() * SP * 4));
            for (size_t u = 0; u < uniq.size(); u++)
                TRY(upload_frames(ctx, frames, (uint64_t)uniq[u], 1, SP, ctx->stage.as<uint32_t>() + u * SP));
            auto slot = [&](int32_t idx) { return (size_t)(std::lower_bound(uniq.begin(), uniq.end(), idx) - uniq.begin()); };
            for (uint32_t k = 0; k < F; k++) {
                sp[k].a = ctx->stage.as<uint32_t>() + slot(src_a[q0 + k]) * SP;
                sp[k].b = src_b[q0 + k] >= 0 ? ctx->stage.as<uint32_t>() + slot(src_b[q0 + k]) * SP : nullptr;
            }
        }
        TRY(ensure(ctx, ctx->srcpairs, F * sizeof(SrcPair)));
        TRY(ensure(ctx, ctx->entries, (size_t)F * P * 2));
        CK(cudaMemcpyAsync(ctx->srcpairs.p, sp.data(), F * sizeof(SrcPair), cudaMemcpyHostToDevice, ctx->st));
        if (ctx->d_map) {   // GBA / NDS: gather through the rescale table
            dim3 qgrid(cdiv(P / 4, 256), F);
            KL(ctx->lc, KC_QUANT, (quantize_k<<<qgrid, 256, 0, ctx->st>>>(ctx->srcpairs.as<SrcPair>(), ctx->d_map, (uint32_t)P, ctx->d_pal, ctx->dual ? 512 : 256,
                                                                          ctx->d_lut, ctx->entries.as<uint16_t>())));
        } else {
            dim3 qgrid(cdiv(P / 8, 256), F);
            if (ctx->lut_full)
                KL(ctx->lc, KC_QUANT, (quantize8_k<true><<<qgrid, 256, 0, ctx->st>>>(ctx->srcpairs.as<SrcPair>(), (uint32_t)P, ctx->d_pal, ctx->dual ? 512 : 256,
                                                                                     ctx->d_lut, ctx->entries.as<uint16_t>())));
            else
                KL(ctx->lc, KC_QUANT, (quantize8_k<false><<<qgrid, 256, 0, ctx->st>>>(ctx->srcpairs.as<SrcPair>(), (uint32_t)P, ctx->d_pal, ctx->dual ? 512 : 256,
                                                                                      ctx->d_lut, ctx->entries.as<uint16_t>())));
        }
        TRY(check_launch(ctx, "quantize"));
        int last_i = -1;
        for (uint32_t k = 0; k < F; k++) {
            uint32_t fc = first_frame_count + q0 + k;
            ep[k].ent = ctx->entries.as<uint16_t>() + (size_t)k * P;
            if (fc % 4 == 0) { ep[k].ient = nullptr; last_i = (int)k; }
            else {
                int ik = (int)k - (int)(fc % 4);
                ep[k].ient = ik >= 0 ? ctx->entries.as<uint16_t>() + (size_t)ik * P : ctx->d_ient;
            }
        }
        TRY(assemble_and_compress(ctx, ep.data(), F, first_frame_count + q0));
        if (last_i >= 0)  // agmv->iframe_entries <- img_entry, src/agmv_encode.c:626-630
            CK(cudaMemcpyAsync(ctx->d_ient, ctx->entries.as<uint16_t>() + (size_t)last_i * P, P * 2, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));  // host vectors and the staging buffer are reused by the next batch
    }
    if (image_bytes) *image_bytes = ctx->image_bytes;
    return OK;
}

extern "C" int agmvb_enc_set_audio_stub(agmvb_ctx* ctx, int on) {
    if (!ctx) return ERR_ARG;
    ctx->lz.stub_bytes = on ? 8u : 0u;
    ctx->lz.audio_chunk = 0;
    ctx->lz.audio = nullptr;
    ctx->lz.audio_size = 0;
    return OK;
}

// ===========================================================================
// audio chunk codec
// ===========================================================================
static int audio_grid(uint64_t n16) { return (int)std::min<uint64_t>(std::max<uint64_t>(cdiv(n16, 256), 1), 148ull * 8); }

// AGMV_CompressAudio (src/agmv_encode.c:659-705): 16-bit samples are companded to one byte each, 8-bit samples are copied
static int audio_compress_dev(agmvb_ctx* ctx, const void* pcm, uint64_t n, int bits) {
    TRY(ensure(ctx, ctx->at_sample, n + 16));
    if (bits == 16) {
        TRY(ensure(ctx, ctx->at_pcm, n * 2 + 32));
        CK(copy_pieces(ctx->at_pcm.p, pcm, n * 2, cudaMemcpyHostToDevice, ctx->st));
        static const uint64_t lut_from = getenv("AGMVB_AUDIO_LUT_FROM") ? strtoull(getenv("AGMVB_AUDIO_LUT_FROM"), nullptr, 10) : (1ull << 20);
        if (n >= lut_from) {  // long track: the 64 KB answer table in shared memory (one persistent CTA pair per SM)
            if (!ctx->at_lut.p) {
                TRY(ensure(ctx, ctx->at_lut, 65536));
                KL(ctx->lc, KC_AUDIO, (audio_lut_k<<<256, 256, 0, ctx->st>>>(ctx->at_lut.as<uint8_t>())));
                CK(cudaFuncSetAttribute(audio_compress16_lut_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            }
            const int grid = (int)std::min<uint64_t>(std::max<uint64_t>(cdiv(n >> 4, AUDIO_LUT_THREADS), 1), 148ull * 2);
            KL(ctx->lc, KC_AUDIO, (audio_compress16_lut_k<<<grid, AUDIO_LUT_THREADS, 65536, ctx->st>>>(ctx->at_pcm.as<uint16_t>(), n, ctx->at_lut.as<uint8_t>(),
                                                                                                       ctx->at_sample.as<uint8_t>())));
        } else {
            KL(ctx->lc, KC_AUDIO, (audio_compress16_k<<<audio_grid(n >> 4), 256, 0, ctx->st>>>(ctx->at_pcm.as<uint16_t>(), n, ctx->at_sample.as<uint8_t>())));
        }
        TRY(check_launch(ctx, "audio_compress16"));
    } else {
        CK(copy_pieces(ctx->at_sample.p, pcm, n, cudaMemcpyHostToDevice, ctx->st));
    }
    return OK;
}

extern "C" int agmvb_enc_set_audio(agmvb_ctx* ctx, const void* pcm, uint64_t audio_size, int bits_per_sample, uint32_t sample_rate,
                                   uint32_t channels, uint32_t total_duration) {
    if (!ctx) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->at_size = 0;
    ctx->lz.audio_chunk = 0; ctx->lz.audio = nullptr; ctx->lz.audio_size = 0;
    if (ctx->lz.stub_bytes > 8) ctx->lz.stub_bytes = 8;
    if (!pcm || audio_size == 0) return OK;
    if (bits_per_sample != 16 && bits_per_sample != 8) FAIL(ERR_ARG, "bits per sample must be 8 or 16");
    TRY(audio_compress_dev(ctx, pcm, audio_size, bits_per_sample));
    CK(cudaStreamSynchronize(ctx->st));  // pcm is the caller's
    ctx->at_size = audio_size; ctx->at_bits = (uint32_t)bits_per_sample; ctx->at_rate = sample_rate; ctx->at_channels = channels;
    ctx->at_duration = total_duration;
    return OK;
}

extern "C" int agmvb_enc_set_audio_chunk(agmvb_ctx* ctx, uint32_t chunk_size) {
    if (!ctx) return ERR_ARG;
    if (!ctx->at_size) FAIL(ERR_ARG, "no audio track: call agmvb_enc_set_audio first");
    ctx->lz.audio = ctx->at_sample.as<uint8_t>();
    ctx->lz.audio_size = ctx->at_size;
    ctx->lz.audio_chunk = chunk_size;
    ctx->lz.stub_bytes = 8 + chunk_size;
    return OK;
}

extern "C" int agmvb_enc_get_atsample(agmvb_ctx* ctx, uint8_t* out, uint64_t cap) {
    if (!ctx || !out) return ERR_ARG;
    if (!ctx->at_size || cap < ctx->at_size) FAIL(ERR_ARG, "no audio track, or buffer smaller than %llu bytes", (unsigned long long)ctx->at_size);
    CK(cudaSetDevice(ctx->device));
    CK(copy_pieces(out, ctx->at_sample.p, ctx->at_size, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

extern "C" int agmvb_audio_compress(agmvb_ctx* ctx, const void* pcm, uint64_t n, int bits_per_sample, uint8_t* atsample) {
    if (!ctx || (n && (!pcm || !atsample)) || (bits_per_sample != 16 && bits_per_sample != 8)) return ERR_ARG;
    if (n == 0) return OK;
    if (ctx->at_size) FAIL(ERR_ARG, "a sequence's audio track is loaded: clear it with agmvb_enc_set_audio(ctx, NULL, 0, ...) first");
    CK(cudaSetDevice(ctx->device));
    TRY(audio_compress_dev(ctx, pcm, n, bits_per_sample));
    CK(copy_pieces(atsample, ctx->at_sample.p, n, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

// the sample loop of AGMV_DecodeAudioChunk (src/agmv_decode.c:431-451)
extern "C" int agmvb_audio_expand(agmvb_ctx* ctx, const uint8_t* atsample, uint64_t n, int bits_per_sample, void* pcm) {
    if (!ctx || (n && (!pcm || !atsample)) || (bits_per_sample != 16 && bits_per_sample != 8)) return ERR_ARG;
    if (n == 0) return OK;
    CK(cudaSetDevice(ctx->device));
    TRY(ensure(ctx, ctx->at_idx, n + 16));
    CK(copy_pieces(ctx->at_idx.p, atsample, n, cudaMemcpyHostToDevice, ctx->st));
    if (bits_per_sample == 16) {
        TRY(ensure(ctx, ctx->d_out, n * 2 + 32));
        KL(ctx->lc, KC_AUDIO, (audio_expand16_k<<<audio_grid(n >> 4), 256, 0, ctx->st>>>(ctx->at_idx.as<uint8_t>(), n, ctx->d_out.as<uint16_t>())));
        TRY(check_launch(ctx, "audio_expand16"));
        CK(copy_pieces(pcm, ctx->d_out.p, n * 2, cudaMemcpyDeviceToHost, ctx->st));
    } else {
        CK(copy_pieces(pcm, ctx->at_idx.p, n, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

// counts[k] = number of grey-equal pixels of coded frames (pair_a[k], pair_b[k]) (AGMV_CompareFrameSimilarity)
extern "C" int agmvb_frame_similarity(agmvb_ctx* ctx, const uint32_t* frames, uint64_t n_frames_in_buffer, int on_device,
                                      const int32_t* pair_a, const int32_t* pair_b, uint32_t n_pairs, uint64_t* counts) {
    if (!ctx || !ctx->enc_ready || !frames || !pair_a || !pair_b || !counts) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (n_pairs == 0) return OK;
    const size_t SP = (size_t)ctx->src_w * ctx->src_h, P = (size_t)ctx->cw * ctx->ch;
    const uint32_t* base = frames;
    if (!on_device) {
        TRY(ensure(ctx, ctx->stage, n_frames_in_buffer * SP * 4));
        TRY(upload_frames(ctx, frames, 0, n_frames_in_buffer, SP, ctx->stage.as<uint32_t>()));
        base = ctx->stage.as<uint32_t>();
    }
    std::vector<SrcPair> sp(n_pairs);
    for (uint32_t k = 0; k < n_pairs; k++) {
        if (pair_a[k] < 0 || pair_b[k] < 0 || (uint64_t)pair_a[k] >= n_frames_in_buffer || (uint64_t)pair_b[k] >= n_frames_in_buffer)
            FAIL(ERR_ARG, "pair %u out of range", k);
        sp[k].a = base + (size_t)pair_a[k] * SP;
        sp[k].b = base + (size_t)pair_b[k] * SP;
    }
    TRY(ensure(ctx, ctx->srcpairs, n_pairs * sizeof(SrcPair)));
    TRY(ensure(ctx, ctx->small, (size_t)n_pairs * 8));
    CK(cudaMemcpyAsync(ctx->srcpairs.p, sp.data(), n_pairs * sizeof(SrcPair), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(ctx->small.p, 0, (size_t)n_pairs * 8, ctx->st));
    for (uint32_t k0 = 0; k0 < n_pairs; k0 += 32768) {
        dim3 grid(std::min<uint32_t>(cdiv(P, 1024), 64), std::min<uint32_t>(32768, n_pairs - k0));
        KL(ctx->lc, KC_MISC, (similarity_k<<<grid, 256, 0, ctx->st>>>(ctx->srcpairs.as<SrcPair>() + k0, ctx->d_map, (uint32_t)P,
                                                                    ctx->small.as<unsigned long long>() + k0)));
    }
    TRY(check_launch(ctx, "similarity"));
    CK(cudaMemcpyAsync(counts, ctx->small.p, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

extern "C" int agmvb_enc_fetch(agmvb_ctx* ctx, uint8_t* image, uint64_t cap, uint32_t* usize, uint32_t* csize) {
    if (!ctx) return ERR_ARG;
    if (image) {
        if (cap < ctx->image_bytes) FAIL(ERR_ARG, "image buffer too small: %llu < %llu", (unsigned long long)cap, (unsigned long long)ctx->image_bytes);
        if (ctx->image_bytes) CK(copy_pieces(image, ctx->image.p, ctx->image_bytes, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    }
    if (usize) memcpy(usize, ctx->last_usize.data(), ctx->last_usize.size() * 4);
    if (csize) memcpy(csize, ctx->last_csize.data(), ctx->last_csize.size() * 4);
    return OK;
}

extern "C" int agmvb_enc_image_ptr(agmvb_ctx* ctx, uint8_t** dev_image, uint64_t* bytes) {
    if (!ctx) return ERR_ARG;
    if (dev_image) *dev_image = ctx->image.as<uint8_t>();
    if (bytes) *bytes = ctx->image_bytes;
    return OK;
}

static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static uint32_t get32(const uint8_t* p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24; }

// AGMV_EncodeHeader, src/agmv_encode.c:21-94 (no audio: CreateAGMV's defaults, src/agmv_utils.c:358-366)
extern "C" int agmvb_enc_header(agmvb_ctx* ctx, uint32_t n_frames, uint32_t fps, uint8_t* out, uint64_t cap, uint64_t* len) {
    if (!ctx || !ctx->enc_ready || !ctx->pal_valid || !out) return ERR_ARG;
    const uint64_t need = 38 + 768 * (ctx->dual ? 2 : 1);
    if (cap < need) FAIL(ERR_ARG, "header needs %llu bytes", (unsigned long long)need);
    memset(out, 0, need);
    memcpy(out, "AGMV", 4);
    put32(out + 4, n_frames);
    put32(out + 8, ctx->cw);
    put32(out + 12, ctx->ch);
    out[16] = 1;
    out[17] = (uint8_t)((ctx->compression == COMP_LZSS ? 0 : 2) + (ctx->dual ? 1 : 2));  // src/agmv_utils.c:487-545
    put32(out + 18, fps);
    out[36] = 16;  // bits per sample
    if (ctx->at_size) {  // the handle carries an audio track (AGMV_WavToAudioTrack, src/agmv_utils.c:1070-1083)
        put32(out + 22, ctx->at_duration);
        put32(out + 26, ctx->at_rate);
        put32(out + 30, (uint32_t)ctx->at_size);
        out[34] = (uint8_t)ctx->at_channels; out[35] = (uint8_t)(ctx->at_channels >> 8);
        out[36] = (uint8_t)ctx->at_bits;
    }
    uint8_t* p = out + 38;
    for (int i = 0; i < (ctx->dual ? 512 : 256); i++) {
        uint32_t c = ctx->h_pal[i];
        *p++ = (uint8_t)(c >> 16); *p++ = (uint8_t)(c >> 8); *p++ = (uint8_t)c;
    }
    if (len) *len = need;
    return OK;
}

enum { SEQ_AGMV = 0, SEQ_VIDEO = 1, SEQ_FULL = 2 };

static int encode_sequence_impl(agmvb_ctx* ctx, int mode, const uint32_t* frames, int on_device, uint32_t n_src, uint32_t w, uint32_t h,
                                uint32_t create_n, uint32_t fps, int opt, int quality, int compression, uint8_t* out, uint64_t cap,
                                uint64_t* out_len, uint32_t* n_encoded) {
    if (!ctx || !frames || !out || n_src < 2) return ERR_ARG;
    TRY(agmvb_enc_begin(ctx, w, h, opt, quality, compression));
    agmvb_enc_set_audio_stub(ctx, mode == SEQ_AGMV);  // only AGMV_EncodeAGMV interleaves (empty) audio chunks
    if (!on_device) {
        // both passes read every frame: upload once when the sequence fits comfortably, else stream it twice
        size_t free_b = 0, total_b = 0;
        const size_t bytes = (size_t)n_src * w * h * 4;
        CK(cudaMemGetInfo(&free_b, &total_b));
        if (bytes + (8ull << 30) < free_b + ctx->seqbuf.cap) {
            TRY(ensure(ctx, ctx->seqbuf, bytes));
            {
                GatedXfer gate(ctx, 0, ctx->st);
                TRY(upload_frames(ctx, frames, 0, n_src, (uint64_t)w * h, ctx->seqbuf.as<uint32_t>()));
                CK(gate.finish(ctx->st));
            }
            frames = ctx->seqbuf.as<uint32_t>();
            on_device = 1;
        }
    }
    TRY(agmvb_enc_histogram(ctx, frames, n_src, on_device));  // every source frame, unscaled (:2371-2568)
    TRY(agmvb_enc_build_palette(ctx));
    // frame numbers are 1-based in the reference: source frame i is frames[i - 1]
    const uint32_t start = 1, end = n_src;
    std::vector<int32_t> sa, sb;
    if (mode == SEQ_FULL) {  // AGMV_EncodeFullAGMV (:4041-4370): every frame as it is
        for (uint32_t i = start; i <= end; i++) { sa.push_back(i - 1); sb.push_back(-1); }
    } else {
        // AGMV_EncodeVideo (:1109-2222) merges a pair only if enough pixels have equal grey values; the ratios of all
        // candidate pairs are measured up front (one launch), the data-dependent walk is then host logic
        std::vector<uint64_t> eq;
        float leniency = 0.f;
        if (mode == SEQ_VIDEO) {
            leniency = opt == OPT_II ? (float)0.1282 : ((opt == OPT_GBA_I || opt == OPT_GBA_II || opt == OPT_GBA_III) ? 0.0f : (float)0.2282);  // :741-790
            std::vector<int32_t> pa(n_src - 1), pb(n_src - 1);
            for (uint32_t k = 0; k + 1 < n_src; k++) { pa[k] = k; pb[k] = k + 1; }
            eq.resize(n_src - 1);
            TRY(agmvb_frame_similarity(ctx, frames, n_src, on_device, pa.data(), pb.data(), n_src - 1, eq.data()));
        }
        const size_t P = (size_t)ctx->cw * ctx->ch;
        // PDIFS schedule (:2727-2770) and loop exit (:3610-3612)
        for (uint32_t i = start; i <= end;) {
            bool merge = true;
            if (mode == SEQ_VIDEO) {
                const uint32_t a = ctx->light ? i + 1 : i;  // 1-based first frame of the pair that would be merged
                if (a + 1 > end) FAIL(ERR_ARG, "sequence too short for AGMV_EncodeVideo's look-ahead");
                merge = (eq[a - 1] / (float)P) >= leniency;
            }
            if (!merge) { sa.push_back(i - 1); sb.push_back(-1); i += 1; }
            else if (ctx->light) {
                if (i + 3 > end) FAIL(ERR_ARG, "sequence too short for the LIGHT schedule");
                sa.push_back(i - 1); sb.push_back(-1);
                sa.push_back(i); sb.push_back(i + 1);
                sa.push_back(i + 2); sb.push_back(-1);
                i += 4;
            } else {
                if (i + 1 > end) FAIL(ERR_ARG, "sequence too short for the HEAVY schedule");
                sa.push_back(i - 1); sb.push_back(i);
                i += 2;
            }
            if (i + 4 >= end) break;
        }
    }
    uint32_t adjusted = end - start;
    switch (opt) {  // :2296-2353
        case OPT_I: case OPT_ANIM: case OPT_GBA_I: case OPT_GBA_II: adjusted /= 2; break;
        case OPT_GBA_III: adjusted = (uint32_t)(adjusted * 0.75f); break;
        default: adjusted = (uint32_t)(adjusted * 0.75); break;
    }
    if (ctx->at_size && mode != SEQ_VIDEO) {
        // audio_chunk->size = audio_size / (f32)frames (:2661-2663 with the adjusted count, :4024-4025 with end - start);
        // AGMV_EncodeFullAGMV writes audio chunks only for handles with a track (:4072-4076)
        const float per_frame = (float)ctx->at_size / (float)(mode == SEQ_AGMV ? adjusted : end - start);
        TRY(agmvb_enc_set_audio_chunk(ctx, (uint32_t)per_frame));
    }
    uint64_t hdr_len = 0, img = 0;
    TRY(agmvb_enc_header(ctx, create_n, fps, out, cap, &hdr_len));
    TRY(agmvb_enc_frames(ctx, frames, n_src, on_device, sa.data(), sb.data(), (uint32_t)sa.size(), 0, &img));
    if (hdr_len + img > cap) FAIL(ERR_ARG, "output buffer too small: need %llu", (unsigned long long)(hdr_len + img));
    TRY(agmvb_enc_fetch(ctx, out + hdr_len, cap - hdr_len, nullptr, nullptr));
    if (mode == SEQ_AGMV) {  // back-patch (:3615-3620)
        put32(out + 4, (uint32_t)sa.size());
        float rate = (float)adjusted / (create_n + 1);
        put32(out + 18, (uint32_t)round(fps * rate));
    } else if (mode == SEQ_VIDEO) {  // :2223-2230
        put32(out + 4, (uint32_t)sa.size());
        float rate = (float)sa.size() / create_n;
        put32(out + 18, (uint32_t)round(fps * rate));
    }
    agmvb_enc_set_audio_stub(ctx, 1);
    ctx->at_size = 0;  // the reference consumes the handle, track included (:3625)
    if (out_len) *out_len = hdr_len + img;
    if (n_encoded) *n_encoded = (uint32_t)sa.size();
    return OK;
}

extern "C" int agmvb_encode_sequence(agmvb_ctx* ctx, const uint32_t* frames, int on_device, uint32_t n_src, uint32_t w, uint32_t h,
                                     uint32_t create_n, uint32_t fps, int opt, int quality, int compression, uint8_t* out, uint64_t cap,
                                     uint64_t* out_len, uint32_t* n_encoded) {
    return encode_sequence_impl(ctx, SEQ_AGMV, frames, on_device, n_src, w, h, create_n, fps, opt, quality, compression, out, cap, out_len, n_encoded);
}
// AGMV_EncodeVideo (src/agmv_encode.c:719-2268): similarity-gated PDIFS; the reference creates its own handle with end-start frames
extern "C" int agmvb_encode_video(agmvb_ctx* ctx, const uint32_t* frames, int on_device, uint32_t n_src, uint32_t w, uint32_t h, uint32_t fps,
                                  int opt, int quality, int compression, uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* n_encoded) {
    return encode_sequence_impl(ctx, SEQ_VIDEO, frames, on_device, n_src, w, h, n_src - 1, fps, opt, quality, compression, out, cap, out_len, n_encoded);
}
// AGMV_EncodeFullAGMV (src/agmv_encode.c:3659-4407): no frame skipping, no audio chunks, header left as CreateAGMV set it
extern "C" int agmvb_encode_full(agmvb_ctx* ctx, const uint32_t* frames, int on_device, uint32_t n_src, uint32_t w, uint32_t h,
                                 uint32_t create_n, uint32_t fps, int opt, int quality, int compression, uint8_t* out, uint64_t cap,
                                 uint64_t* out_len, uint32_t* n_encoded) {
    return encode_sequence_impl(ctx, SEQ_FULL, frames, on_device, n_src, w, h, create_n, fps, opt, quality, compression, out, cap, out_len, n_encoded);
}

// ===========================================================================
// one sequence over several GPUs of this process
// ===========================================================================
// AGMV_EncodeAGMV shards by frame range once the palette is known (P-frames only refer to the I-frame of their own group
// of four encoded frames): every GPU histograms its share of the SOURCE frames, the bins are summed (the palette is a
// function of every frame, src/agmv_encode.c:2371-2568), every GPU builds the identical palette and encodes a range of the
// schedule whose borders are whole PDIFS groups and whole GOPs (12 encoded frames for the LIGHT profiles, 4 for HEAVY), and
// the chunk images are copied straight to their final offsets of the caller's buffer. One host thread per GPU; the only
// exchanges are the bins (peer copies over NVLink, summed on the first device) and the image sizes (host).
__global__ void hist_sum_k(unsigned long long* __restrict__ acc, const unsigned long long* __restrict__ part, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] += part[i];
}

namespace {
struct HostBarrier {   // reusable barrier for the per-GPU host threads
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0, phase = 0;
    explicit HostBarrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const int ph = phase;
        if (++waiting == n) { waiting = 0; phase++; cv.notify_all(); }
        else cv.wait(lk, [&] { return phase != ph; });
    }
};
}  // namespace

// the AGMV_EncodeAGMV schedule (src/agmv_encode.c:2727-2770, exit rule :3610-3612) for source frames 1..n_src
static void agmv_schedule(bool light, uint32_t n_src, std::vector<int32_t>& sa, std::vector<int32_t>& sb) {
    sa.clear(); sb.clear();
    for (uint32_t i = 1; i <= n_src;) {
        if (light) {
            sa.push_back(i - 1); sb.push_back(-1);
            sa.push_back(i); sb.push_back(i + 1);
            sa.push_back(i + 2); sb.push_back(-1);
            i += 4;
        } else {
            sa.push_back(i - 1); sb.push_back(i);
            i += 2;
        }
        if (i + 4 >= n_src) break;
    }
}

extern "C" int agmvb_shard_range(uint32_t n_enc, int light, int n_shards, int shard, uint32_t* first, uint32_t* count) {
    if (n_shards <= 0 || shard < 0 || shard >= n_shards || !first || !count) return ERR_ARG;
    const uint32_t group = light ? 12u : 4u;   // whole PDIFS groups and whole GOPs
    const uint32_t units = (n_enc + group - 1) / group, per = (units + n_shards - 1) / n_shards;
    const uint32_t a = std::min<uint64_t>((uint64_t)shard * per * group, n_enc), b = std::min<uint64_t>((uint64_t)(shard + 1) * per * group, n_enc);
    *first = a;
    *count = b - a;
    return OK;
}

extern "C" int agmvb_encode_sequence_multi(agmvb_ctx* const* ctxs, int n_ctx, const uint32_t* frames, uint32_t n_src, uint32_t w, uint32_t h,
                                           uint32_t create_n, uint32_t fps, int opt, int quality, int compression, uint8_t* out, uint64_t cap,
                                           uint64_t* out_len, uint32_t* n_encoded) {
    if (!ctxs || n_ctx < 1 || !frames || !out || n_src < 2) return ERR_ARG;
    for (int k = 0; k < n_ctx; k++) if (!ctxs[k]) return ERR_ARG;
    agmvb_ctx* ctx = ctxs[0];   // error text of argument checks
    if (compression != COMP_LZSS) FAIL(ERR_UNSUPPORTED, "LZ77 carries its bitstream buffer from frame to frame: not sharded");
    if (n_ctx == 1) return agmvb_encode_sequence(ctx, frames, 0, n_src, w, h, create_n, fps, opt, quality, compression, out, cap, out_len, n_encoded);
    const bool light = opt_is_light(opt);
    std::vector<int32_t> sa, sb;
    if (light ? n_src < 4 : n_src < 2) FAIL(ERR_ARG, "sequence too short");
    agmv_schedule(light, n_src, sa, sb);
    const uint32_t n_enc = (uint32_t)sa.size();
    std::vector<int> rc(n_ctx, OK);
    std::vector<uint64_t> img(n_ctx, 0);
    uint64_t hdr_len = 0;
    HostBarrier bar(n_ctx);
    const uint64_t fpx = (uint64_t)w * h;
    auto fail_all = [&](int k, int code) { rc[k] = code; };
    auto any_failed = [&]() { for (int c : rc) if (c != OK) return true; return false; };
    auto worker = [&](int k) {
        agmvb_ctx* c = ctxs[k];
        int r = agmvb_enc_begin(c, w, h, opt, quality, compression);
        if (!r) r = agmvb_enc_set_audio_stub(c, 1);
        // pass 1: this GPU's share of the source frames
        const uint64_t f0 = (uint64_t)n_src * k / n_ctx, f1 = (uint64_t)n_src * (k + 1) / n_ctx;
        if (!r && f1 > f0) r = agmvb_enc_histogram(c, frames + f0 * fpx, f1 - f0, 0);
        if (!r && cudaStreamSynchronize(c->st) != cudaSuccess) r = ERR_CUDA;
        if (r) fail_all(k, r);
        bar.wait();
        if (k == 0 && !any_failed()) {   // sum on the first device, hand the total back to everybody
            agmvb_ctx* c0 = ctxs[0];
            const uint32_t bins = c0->mc + 1;
            unsigned long long* tmp = c0->d_keys[0];   // free until the palette sort
            cudaError_t e = cudaSetDevice(c0->device);
            for (int j = 1; j < n_ctx; j++) {   // direct NVLink path where the topology has one (else the copies are staged by the driver)
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, c0->device, ctxs[j]->device) == cudaSuccess && can)
                    if (cudaDeviceEnablePeerAccess(ctxs[j]->device, 0) != cudaSuccess) (void)cudaGetLastError();   // already enabled
            }
            for (int j = 1; j < n_ctx && e == cudaSuccess; j++) {
                e = cudaMemcpyPeerAsync(tmp, c0->device, ctxs[j]->d_hist, ctxs[j]->device, (size_t)bins * 8, c0->st);
                if (e == cudaSuccess) hist_sum_k<<<cdiv(bins, 256), 256, 0, c0->st>>>(c0->d_hist, tmp, bins);
            }
            for (int j = 1; j < n_ctx && e == cudaSuccess; j++)
                e = cudaMemcpyPeerAsync(ctxs[j]->d_hist, ctxs[j]->device, c0->d_hist, c0->device, (size_t)bins * 8, c0->st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c0->st);
            if (e != cudaSuccess) { snprintf(c0->err, sizeof c0->err, "histogram exchange: %s", cudaGetErrorString(e)); fail_all(0, ERR_CUDA); }
        }
        bar.wait();
        r = any_failed() ? ERR_CUDA : OK;
        if (!r) r = agmvb_enc_build_palette(c);   // identical on every GPU
        uint32_t e0 = 0, cnt = 0;
        agmvb_shard_range(n_enc, light, n_ctx, k, &e0, &cnt);
        if (!r && k == 0) r = agmvb_enc_header(c, create_n, fps, out, cap, &hdr_len);
        if (!r && cnt) r = agmvb_enc_frames(c, frames, n_src, 0, sa.data() + e0, sb.data() + e0, cnt, e0, &img[k]);
        if (r) fail_all(k, r);
        bar.wait();
        if (any_failed()) return;
        uint64_t off = hdr_len, total = hdr_len;
        for (int j = 0; j < n_ctx; j++) { if (j < k) off += img[j]; total += img[j]; }
        if (total > cap) { if (k == 0) snprintf(c->err, sizeof c->err, "output buffer too small: need %llu", (unsigned long long)total); fail_all(k, ERR_ARG); return; }
        if (cnt) r = agmvb_enc_fetch(c, out + off, img[k], nullptr, nullptr);   // straight to its place in the file
        if (r) fail_all(k, r);
    };
    std::vector<std::thread> th;
    for (int k = 1; k < n_ctx; k++) th.emplace_back(worker, k);
    worker(0);
    for (auto& t : th) t.join();
    for (int k = 0; k < n_ctx; k++)
        if (rc[k] != OK) {
            if (k != 0) snprintf(ctx->err, sizeof ctx->err, "GPU %d: %.400s", ctxs[k]->device, ctxs[k]->err);
            return rc[k];
        }
    uint64_t total = hdr_len;
    for (int k = 0; k < n_ctx; k++) total += img[k];
    // back-patch (:3615-3620)
    uint32_t adjusted = n_src - 1;
    switch (opt) {
        case OPT_I: case OPT_ANIM: case OPT_GBA_I: case OPT_GBA_II: adjusted /= 2; break;
        case OPT_GBA_III: adjusted = (uint32_t)(adjusted * 0.75f); break;
        default: adjusted = (uint32_t)(adjusted * 0.75); break;
    }
    put32(out + 4, n_enc);
    const float rate = (float)adjusted / (create_n + 1);
    put32(out + 18, (uint32_t)round(fps * rate));
    if (out_len) *out_len = total;
    if (n_encoded) *n_encoded = n_enc;
    return OK;
}

// ===========================================================================
// decoder
// ===========================================================================
// occurrences of a four-byte tag at any byte offset; sixteen offsets per thread from five aligned words (the file image is
// 16-byte aligned and followed by at least 64 readable zero bytes)
__global__ void __launch_bounds__(256) count_fourcc_k(const uint8_t* __restrict__ d, uint64_t len, uint32_t fourcc, unsigned long long* __restrict__ count) {
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16u;
    uint32_t hits = 0;
    if (i0 < len) {
        const uint4 a = *reinterpret_cast<const uint4*>(d + i0);
        const uint32_t w[5] = {a.x, a.y, a.z, a.w, *reinterpret_cast<const uint32_t*>(d + i0 + 16)};
#pragma unroll
        for (int k = 0; k < 16; k++)
            hits += __funnelshift_r(w[k >> 2], w[(k >> 2) + 1], (k & 3) * 8) == fourcc && i0 + k + 4 <= len;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
    if (hits && lane_id() == 0) atomicAdd(count, (unsigned long long)hits);
}

// Walk the file the way AGMV_DecodeAGMV does: AGMV_FindNextFrameChunk (src/agmv_utils.c:140-166) checks the
// four bytes at the cursor, then every offset from cursor+4 on. `cursor` after a frame is wherever the bit
// reader stopped; when no stray 'AGFC' hides in a payload (verified on the device by counting occurrences)
// any cursor inside the payload finds the same next chunk, so the index can be built before decoding.
static int64_t find_next_agfc(const uint8_t* f, uint64_t len, uint64_t p) {
    if (p + 4 <= len && !memcmp(f + p, "AGFC", 4)) return (int64_t)p;
    for (uint64_t q = p + 4; q + 4 <= len; q++) {
        const uint8_t* a = (const uint8_t*)memchr(f + q, 'A', len - 3 - q);
        if (!a) return -1;
        q = (uint64_t)(a - f);
        if (!memcmp(a, "AGFC", 4)) return (int64_t)q;
    }
    return -1;
}

// exact number of payload bytes the reference's bit reader touches for one LZSS chunk (host, tokens only)
static uint64_t lzss_consumed_host(const uint8_t* f, uint64_t len, uint64_t data_off, uint32_t usize, uint32_t csize) {
    uint64_t rp = data_off, bits = 0, nbits = (uint64_t)csize * 8, bpos = 0;
    uint32_t acc = 0, navail = 0;
    auto rd = [&](uint32_t nb) {
        while (navail < nb) { uint32_t b = rp < len ? f[rp] : 0; rp++; acc |= b << navail; navail += 8; }
        uint32_t v = acc & ((1u << nb) - 1); acc >>= nb; navail -= nb; return v;
    };
    while (bits < nbits && bpos < usize) {
        uint32_t flag = rd(1); bits++;
        if (flag) { rd(8); bits += 8; bpos++; }
        else {
            uint64_t off = rd(16); uint32_t l = rd(4); bits += 20;
            if (off >= 1 && off <= bpos) bpos += l;                                   // ordinary match
            else if (off > bpos && off < bpos + l && bpos > 0) bpos += l - (off - bpos);  // unsigned wrap: the tail copies from index 0
        }
    }
    return rp - data_off;
}

static void park_stream(agmvb_ctx* ctx, DecStream& s) {
    DecStream k;
    k.d_file = s.d_file; k.file_cap = s.file_cap; k.d_pal = s.d_pal; k.d_img = s.d_img; k.d_ifr = s.d_ifr; k.d_persist = s.d_persist;
    k.persist_len = s.persist_len; k.d_ring = s.d_ring; k.w = s.w; k.h = s.h;
    cudaFree(s.d_snap);
    // bounded: the oldest parked buffers go first (AGMVB_PARK_MB per context, default a third of the device's memory; 0 = park nothing)
    static const uint64_t park_cap = [] {
        if (getenv("AGMVB_PARK_MB")) return strtoull(getenv("AGMVB_PARK_MB"), nullptr, 10) << 20;
        size_t free_b = 0, total_b = 0;
        return cudaMemGetInfo(&free_b, &total_b) == cudaSuccess ? (unsigned long long)(total_b / 3) : (8192ull << 20);
    }();
    auto bytes_of = [](const DecStream& c) { return c.file_cap + (uint64_t)c.w * c.h * 8 + c.persist_len + (c.d_ring ? (uint64_t)DEC_RING * c.w * c.h * 4 : 0) + 2048; };
    ctx->parked.push_back(k);
    uint64_t total = 0;
    for (const DecStream& c : ctx->parked) total += bytes_of(c);
    while (!ctx->parked.empty() && (total > park_cap || ctx->parked.size() > 1024)) {
        total -= bytes_of(ctx->parked.front());
        free_stream(ctx->parked.front());
        ctx->parked.erase(ctx->parked.begin());
    }
    s = DecStream();
}
static bool unpark_stream(agmvb_ctx* ctx, DecStream& s, uint64_t file_bytes, size_t P) {
    for (size_t k = 0; k < ctx->parked.size(); k++) {
        DecStream& c = ctx->parked[k];
        if ((size_t)c.w * c.h == P && c.file_cap >= file_bytes + 64) {
            s.d_file = c.d_file; s.file_cap = c.file_cap; s.d_pal = c.d_pal; s.d_img = c.d_img; s.d_ifr = c.d_ifr; s.d_persist = c.d_persist;
            s.persist_len = c.persist_len; s.d_ring = c.d_ring;
            ctx->parked.erase(ctx->parked.begin() + k);
            return true;
        }
    }
    return false;
}

extern "C" int agmvb_dec_open(agmvb_ctx* ctx, const uint8_t* file, uint64_t len, int* stream, uint32_t* w, uint32_t* h, uint32_t* n_frames) {
    if (!ctx || !file || !stream) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    // AGMV_DecodeHeader, src/agmv_decode.c:91-143
    if (len < 38 || memcmp(file, "AGMV", 4)) FAIL(ERR_HEADER, "not an AGMV stream");
    const uint32_t nfr = get32(file + 4), W = get32(file + 8), H = get32(file + 12);
    const int version = file[17];
    const uint32_t fps = get32(file + 18);
    const int bps = file[36] | file[37] << 8;
    if (!(version >= 1 && version <= 4) || fps >= 200 || !(bps == 16 || bps == 8)) FAIL(ERR_HEADER, "invalid header fields");
    if (W == 0 || H == 0 || (W & 3) || (H & 3)) FAIL(ERR_UNSUPPORTED, "width and height must be multiples of 4");
    const int dual = version == 1 || version == 3;
    const uint64_t pal_bytes = 768ull * (dual ? 2 : 1);
    if (len < 38 + pal_bytes) FAIL(ERR_HEADER, "truncated palette");
    uint32_t pal[512];
    memset(pal, 0, sizeof pal);
    for (int i = 0; i < (dual ? 512 : 256); i++) {
        const uint8_t* p = file + 38 + 3 * i;
        pal[i] = (uint32_t)p[0] << 16 | (uint32_t)p[1] << 8 | p[2];  // AGIDL_RGB(r,g,b,RGB_888)
    }
    // W and H come from the file: bound what a damaged header can make this call allocate (AGMVB_MAX_PIXELS, default 2^28)
    static const uint64_t max_px = getenv("AGMVB_MAX_PIXELS") ? strtoull(getenv("AGMVB_MAX_PIXELS"), nullptr, 10) : (1ull << 28);
    if ((uint64_t)W * H > max_px) FAIL(ERR_UNSUPPORTED, "%u x %u pixels exceed AGMVB_MAX_PIXELS", W, H);
    DecStream s;
    struct Guard {   // every error return below releases what the stream holds on the device
        DecStream* s;
        ~Guard() { if (s) free_stream(*s); }
    } guard{&s};
    s.w = W; s.h = H; s.n_frames = nfr; s.version = (uint32_t)version; s.dual = dual; s.lz77 = version >= 3; s.file_len = len;
    const uint32_t audio_duration = get32(file + 22);
    const size_t P = (size_t)W * H;
    // chunk index
    uint64_t cursor = 38 + pal_bytes;
    for (uint32_t i = 0; i < nfr; i++) {
        int64_t at = find_next_agfc(file, len, cursor);
        if (at < 0 || (uint64_t)at + 16 > len) FAIL(ERR_HEADER, "frame chunk %u not found", i);
        uint32_t us = get32(file + at + 8), cs = get32(file + at + 12);
        if (us > 2 * P + 64)
            FAIL(ERR_MEMORY, "frame %u: uncompressed size %u exceeds the reference's bitstream buffer", i, us);
        s.data_off.push_back((uint64_t)at + 16);
        s.usize.push_back(us);
        s.csize.push_back(cs);
        cursor = (uint64_t)at + 16 + cs;
        if (audio_duration != 0) {
            // AGMV_FindNextAudioChunk + AGMV_DecodeAudioChunk consume 'AGAC', size, payload (src/agmv_decode.c:412-453)
            uint64_t q = cursor;
            int64_t aa = -1;
            if (q + 4 <= len && !memcmp(file + q, "AGAC", 4)) aa = (int64_t)q;
            else for (uint64_t r = q + 4; r + 4 <= len; r++) if (!memcmp(file + r, "AGAC", 4)) { aa = (int64_t)r; break; }
            if (aa >= 0 && (uint64_t)aa + 8 <= len) { cursor = (uint64_t)aa + 8 + get32(file + aa + 4); s.audio_off.push_back((uint64_t)aa + 8); s.audio_len.push_back(get32(file + aa + 4)); }
        }
    }
    s.audio_bits = (uint32_t)bps;
    if (!unpark_stream(ctx, s, len, P)) {
        s.file_cap = len + len / 8 + 4096;
        CK(cudaMalloc(&s.d_file, s.file_cap));
        CK(cudaMalloc(&s.d_pal, 512 * 4));
        CK(cudaMalloc(&s.d_img, P * 4));
        CK(cudaMalloc(&s.d_ifr, P * 4));
        s.persist_len = (uint32_t)(2 * P + 64);
        CK(cudaMalloc(&s.d_persist, s.persist_len));
    }
    CK(copy_pieces(s.d_file, file, len, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(s.d_file + len, 0, 64, ctx->st));
    CK(cudaMemcpyAsync(s.d_pal, pal, 512 * 4, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(s.d_img, 0, P * 4, ctx->st));      // defined start state: zero pages
    CK(cudaMemsetAsync(s.d_ifr, 0, P * 4, ctx->st));
    CK(cudaMemsetAsync(s.d_persist, 0, s.persist_len, ctx->st));
    // stray 'AGFC' inside a payload would make the chunk walk depend on where the bit reader stopped
    TRY(ensure(ctx, ctx->d_count, 8));
    CK(cudaMemsetAsync(ctx->d_count.p, 0, 8, ctx->st));
    KL(ctx->lc, KC_MISC, (count_fourcc_k<<<cdiv(cdiv(len, 16), 256), 256, 0, ctx->st>>>(s.d_file, len, 0x43464741u, ctx->d_count.as<unsigned long long>())));
    unsigned long long hits = 0;
    CK(cudaMemcpyAsync(&hits, ctx->d_count.p, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (hits != nfr && !s.lz77) {
        // rare slow path: replay the token stream on the host to learn the exact cursor after every frame
        s.data_off.clear(); s.usize.clear(); s.csize.clear(); s.audio_off.clear(); s.audio_len.clear();
        cursor = 38 + pal_bytes;
        for (uint32_t i = 0; i < nfr; i++) {
            int64_t at = find_next_agfc(file, len, cursor);
            if (at < 0 || (uint64_t)at + 16 > len) FAIL(ERR_HEADER, "frame chunk %u not found", i);
            uint32_t us = get32(file + at + 8), cs = get32(file + at + 12);
            if (us > 2 * P + 64) FAIL(ERR_MEMORY, "frame %u too large", i);
            s.data_off.push_back((uint64_t)at + 16); s.usize.push_back(us); s.csize.push_back(cs);
            cursor = (uint64_t)at + 16 + lzss_consumed_host(file, len, (uint64_t)at + 16, us, cs);
            if (audio_duration != 0) {
                int64_t aa = -1;
                if (cursor + 4 <= len && !memcmp(file + cursor, "AGAC", 4)) aa = (int64_t)cursor;
                else for (uint64_t r = cursor + 4; r + 4 <= len; r++) if (!memcmp(file + r, "AGAC", 4)) { aa = (int64_t)r; break; }
                if (aa >= 0 && (uint64_t)aa + 8 <= len) { cursor = (uint64_t)aa + 8 + get32(file + aa + 4); s.audio_off.push_back((uint64_t)aa + 8); s.audio_len.push_back(get32(file + aa + 4)); }
            }
        }
    }
    s.open = true;
    int id = -1;
    for (size_t k = 0; k < ctx->streams.size(); k++) if (!ctx->streams[k].open) { id = (int)k; break; }
    if (id < 0) { ctx->streams.push_back(DecStream()); id = (int)ctx->streams.size() - 1; }
    ctx->streams[id] = s;
    guard.s = nullptr;
    *stream = id;
    if (w) *w = W;
    if (h) *h = H;
    if (n_frames) *n_frames = nfr;
    return OK;
}

// The audio half of AGMV_DecodeAGMV's loop (src/agmv_decode.c:572-587): every 'AGAC' chunk that follows a frame chunk, in
// order, expanded (16-bit tracks) or copied (8-bit) to audio_track->pcm / pcm8 at the running start_point. One launch for
// the whole track; the chunk positions come from the index agmvb_dec_open built.
extern "C" int agmvb_dec_audio(agmvb_ctx* ctx, int stream, void* pcm, uint64_t cap_samples, uint64_t* n_samples, int* bits_per_sample) {
    if (!ctx || stream < 0 || (size_t)stream >= ctx->streams.size() || !ctx->streams[stream].open || ctx->streams[stream].raw) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    DecStream& d = ctx->streams[stream];
    const size_t nc = d.audio_off.size();
    uint64_t total = 0;
    uint32_t maxlen = 0;
    std::vector<uint64_t> dst(nc);
    for (size_t c = 0; c < nc; c++) { dst[c] = total; total += d.audio_len[c]; maxlen = std::max(maxlen, d.audio_len[c]); }
    if (n_samples) *n_samples = total;
    if (bits_per_sample) *bits_per_sample = (int)d.audio_bits;
    if (!pcm || total == 0) return OK;  // size query, or a stream without a track
    if (cap_samples < total) FAIL(ERR_ARG, "audio track has %llu samples", (unsigned long long)total);
    const size_t bytes = total * (d.audio_bits == 16 ? 2 : 1);
    TRY(ensure(ctx, ctx->d_out, bytes + 32));
    TRY(ensure(ctx, ctx->at_idx, nc * 20 + 64));
    uint8_t* base = ctx->at_idx.as<uint8_t>();
    uint64_t* d_off = reinterpret_cast<uint64_t*>(base);
    uint64_t* d_dst = d_off + nc;
    uint32_t* d_len = reinterpret_cast<uint32_t*>(d_dst + nc);
    CK(cudaMemcpyAsync(d_off, d.audio_off.data(), nc * 8, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(d_dst, dst.data(), nc * 8, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(d_len, d.audio_len.data(), nc * 4, cudaMemcpyHostToDevice, ctx->st));
    for (size_t c0 = 0; c0 < nc; c0 += 32768) {
        dim3 grid(std::min<uint32_t>(std::max<uint32_t>(cdiv(maxlen, 256), 1), 64), (uint32_t)std::min<size_t>(32768, nc - c0));
        KL(ctx->lc, KC_AUDIO, (audio_track_k<<<grid, 256, 0, ctx->st>>>(d.d_file, d.file_len, d_off + c0, d_len + c0, d_dst + c0, (int)d.audio_bits,
                                                                      ctx->d_out.as<uint16_t>(), ctx->d_out.as<uint8_t>())));
    }
    TRY(check_launch(ctx, "audio_track"));
    CK(copy_pieces(pcm, ctx->d_out.p, bytes, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));  // dst / the index vectors are host memory of this call
    return OK;
}

extern "C" int agmvb_dec_close(agmvb_ctx* ctx, int stream) {
    if (!ctx || stream < 0 || (size_t)stream >= ctx->streams.size() || !ctx->streams[stream].open) return ERR_ARG;
    CK(cudaStreamSynchronize(ctx->st));
    if (ctx->streams[stream].raw) free_stream(ctx->streams[stream]); else park_stream(ctx, ctx->streams[stream]);
    return OK;
}

// AGMV_SkipTo (src/agmv_playback.c:94-100) without the rounding to an I-frame: the next decoded frame is `frame_index`,
// decoded with frame_count = frame_index; pixels, I-frame snapshot and bitstream leftovers stay what the last decoded
// frame made them, exactly as in the reference, whose seek only moves the file cursor and sets frame_count.
extern "C" int agmvb_dec_seek(agmvb_ctx* ctx, int stream, uint32_t frame_index) {
    if (!ctx || stream < 0 || (size_t)stream >= ctx->streams.size() || !ctx->streams[stream].open) return ERR_ARG;
    DecStream& d = ctx->streams[stream];
    if (d.raw) FAIL(ERR_ARG, "chunk-fed streams are positioned by the caller (frame_count argument of agmvb_dec_chunk)");
    if (frame_index >= d.n_frames) FAIL(ERR_ARG, "frame %u of %u", frame_index, d.n_frames);
    d.next = frame_index;
    return OK;
}


// Decode the next `count` frames of each listed stream (all of one size).
// outs[s]: device destination (count*P pixels) or nullptr for the per-stream ring.
static int dec_batch_impl(agmvb_ctx* ctx, const int* ids, uint32_t S, uint32_t count, uint32_t* const* outs, uint64_t* cks) {
    if (S == 0 || count == 0) return OK;
    for (uint32_t s = 0; s < S; s++) {
        if (ids[s] < 0 || (size_t)ids[s] >= ctx->streams.size() || !ctx->streams[ids[s]].open) FAIL(ERR_ARG, "bad stream handle");
        DecStream& d = ctx->streams[ids[s]];
        if (d.w != ctx->streams[ids[0]].w || d.h != ctx->streams[ids[0]].h) FAIL(ERR_ARG, "streams of one batch must share a frame size");
        if (!d.raw && d.next + count > d.n_frames) FAIL(ERR_ARG, "stream %d has only %u frames left", ids[s], d.n_frames - d.next);
    }
    const uint32_t W = ctx->streams[ids[0]].w, H = ctx->streams[ids[0]].h, B = (W >> 2) * (H >> 2);
    const size_t P = (size_t)W * H;
    for (uint32_t s = 0; s < S; s++) {
        DecStream& d = ctx->streams[ids[s]];
        if (!outs[s] && !d.d_ring) { CK(cudaMalloc(&d.d_ring, (size_t)DEC_RING * P * 4)); }
    }
    if (cks) {
        TRY(ensure(ctx, ctx->d_cksum, (size_t)S * count * 8));
        CK(cudaMemsetAsync(ctx->d_cksum.p, 0, (size_t)S * count * 8, ctx->st));
    }
    // frames per stream per chunk: bound the expansion + record workspace (~3 GB)
    uint64_t worst = 0;
    for (uint32_t s = 0; s < S; s++) {
        DecStream& d = ctx->streams[ids[s]];
        for (uint32_t k = 0; k < count; k++) {
            uint64_t e = d.lz77 ? std::min<uint64_t>((uint64_t)(d.csize[d.at(d.next + k)] / 4 + 1) * 256, d.persist_len) : d.usize[d.at(d.next + k)];
            worst = std::max<uint64_t>(worst, e + DEC_SLACK);
        }
    }
    const uint64_t per_frame = worst + (uint64_t)B * 4;
    uint32_t C = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(count, (3ull << 30) / (per_frame * S)));
    if (C < count && C > 4) C &= ~3u;

    std::vector<DecFrame> fr;
    std::vector<DecStep> steps;
    std::vector<RecMeta> rmeta;
    std::vector<uint32_t*> rdst;
    std::vector<RecStream> rstreams;
    std::vector<uint8_t> stepbuf;
    for (uint32_t c0 = 0; c0 < count; c0 += C) {
        const uint32_t cn = std::min(C, count - c0);
        const uint32_t F = S * cn;
        fr.resize(F);
        uint64_t eoff = 0;
        for (uint32_t s = 0; s < S; s++) {
            DecStream& d = ctx->streams[ids[s]];
            for (uint32_t k = 0; k < cn; k++) {
                DecFrame& x = fr[s * cn + k];
                uint32_t g = d.next + k;
                x.file = d.d_file; x.file_len = d.file_len; x.data_off = d.data_off[d.at(g)]; x.ebuf_off = eoff;
                x.persist = d.d_persist; x.persist_len = d.persist_len; x.usize = d.usize[d.at(g)]; x.csize = d.csize[d.at(g)];
                x.stream_first = s * cn; x.lz77 = d.lz77; x.dual = d.dual;
                uint64_t e = d.lz77 ? std::min<uint64_t>((uint64_t)(d.csize[d.at(g)] / 4 + 1) * 256, d.persist_len) : d.usize[d.at(g)];
                eoff += (e + DEC_SLACK + 15) & ~15ull;
            }
        }
        TRY(ensure(ctx, ctx->d_frames, F * sizeof(DecFrame)));
        TRY(ensure(ctx, ctx->d_ebuf, eoff + 64));
        TRY(ensure(ctx, ctx->d_bpos, (size_t)F * 4));
        TRY(ensure(ctx, ctx->d_consumed, (size_t)F * 4));
        TRY(ensure(ctx, ctx->d_stale, (size_t)F * 4));
        TRY(ensure(ctx, ctx->d_recs, (size_t)F * B * 4));
        CK(cudaMemcpyAsync(ctx->d_frames.p, fr.data(), F * sizeof(DecFrame), cudaMemcpyHostToDevice, ctx->st));
        const DecFrame* dfr = ctx->d_frames.as<DecFrame>();
        // block-index segments: one per frame over the safe prefix of its expansion
        std::vector<OrbitSeg> segs(F);
        uint32_t ntile = 0, max_tiles = 1;
        for (uint32_t k = 0; k < F; k++) {
            const uint64_t next_off = k + 1 < F ? fr[k + 1].ebuf_off : eoff;
            segs[k].off = fr[k].ebuf_off;
            segs[k].cap_len = (uint32_t)(next_off - fr[k].ebuf_off);
            segs[k].tile_base = ntile;
            ntile += orbit_tiles(segs[k].cap_len);
            max_tiles = std::max(max_tiles, orbit_tiles(segs[k].cap_len));
        }
        TRY(ensure(ctx, ctx->d_code, eoff + 64));
        TRY(ensure(ctx, ctx->d_segs, F * sizeof(OrbitSeg)));
        TRY(ensure(ctx, ctx->d_seglen, (size_t)F * 4));
        TRY(ensure(ctx, ctx->d_oexit, (size_t)ntile * ORB_SP + 64));
        TRY(ensure(ctx, ctx->d_ow, (size_t)ntile * ORB_SP * 2 + 64));
        TRY(ensure(ctx, ctx->d_oentry, (size_t)ntile + 64));
        TRY(ensure(ctx, ctx->d_ocum, (size_t)ntile * 4 + 64));
        TRY(ensure(ctx, ctx->d_ofinal, (size_t)F * 8 + 64));
        CK(cudaMemcpyAsync(ctx->d_segs.p, segs.data(), F * sizeof(OrbitSeg), cudaMemcpyHostToDevice, ctx->st));
        OrbitTables tb;
        tb.exit_tab = ctx->d_oexit.as<uint8_t>(); tb.w_tab = ctx->d_ow.as<uint16_t>(); tb.entry_tab = ctx->d_oentry.as<uint8_t>();
        tb.cumbase = ctx->d_ocum.as<uint32_t>(); tb.final_pos = ctx->d_ofinal.as<uint32_t>(); tb.final_cum = tb.final_pos + F;
        // Frame-parallel reconstruction of the frames between snapshots (see below) when every frame goes to its own buffer, the
        // picture is more than one block wide and no checksums are wanted; AGMVB_RECON_SERIAL=1 keeps everything in order.
        static const bool serial_only = getenv("AGMVB_RECON_SERIAL") && atoi(getenv("AGMVB_RECON_SERIAL")) != 0;
        bool par = !serial_only && (W >> 2) > 1 && cn > 1 && !cks;
        for (uint32_t s = 0; s < S; s++) par = par && outs[s] != nullptr;
        if (par) {
            TRY(ensure(ctx, ctx->d_keeps, (size_t)S * B));
            CK(cudaMemsetAsync(ctx->d_keeps.p, 0, (size_t)S * B, ctx->st));
        }
        KL(ctx->lc, KC_EXPAND, (expand_mrr_k<<<cdiv(F, EX_WARPS), EX_WARPS * 32, 0, ctx->st>>>(dfr, F, ctx->d_ebuf.as<uint8_t>(), ctx->d_bpos.as<uint32_t>(),
                                                                                            ctx->d_consumed.as<uint32_t>())));
        KL(ctx->lc, KC_STALE, (stale_k<<<cdiv(F, 64), 64, 0, ctx->st>>>(dfr, ctx->d_bpos.as<uint32_t>(), F, ctx->d_ebuf.as<uint8_t>(), ctx->d_stale.as<uint8_t>())));
        {
            dim3 sgrid(std::min<uint32_t>(max_tiles * 4, 1024), F);
            KL(ctx->lc, KC_INDEX, (index_steps_k<<<sgrid, 256, 0, ctx->st>>>(dfr, ctx->d_bpos.as<uint32_t>(), ctx->d_ebuf.as<uint8_t>(), ctx->d_code.as<uint8_t>(),
                                                                            ctx->d_seglen.as<uint32_t>())));
        }
        orbit_run<33, IndexStep>(ctx->d_code.as<uint8_t>(), ctx->d_segs.as<OrbitSeg>(), F, ctx->d_seglen.as<uint32_t>(), ntile, tb,
                                 IndexVisit{ctx->d_recs.as<uint32_t>(), B}, ctx->lc, KC_INDEX);
        KL(ctx->lc, KC_INDEX, (index_tail_k<<<cdiv(F, 64), 64, 0, ctx->st>>>(dfr, ctx->d_bpos.as<uint32_t>(), F, ctx->d_ebuf.as<uint8_t>(), ctx->d_stale.as<uint8_t>(), B,
                                                                            tb.final_pos, tb.final_cum, ctx->d_recs.as<uint32_t>(), par ? ctx->d_keeps.as<uint8_t>() : nullptr, cn)));
        TRY(check_launch(ctx, "expand/index"));
        // reconstruction, frame by frame; step layout [k][s]
        // Reconstruction. Layout of the per-frame arrays: [k][s]. A block whose pixels can depend on the previous frame (it keeps
        // them in some frame of the chunk: d_keeps; or it is the last block) is walked through every frame in order; every other
        // block only through the snapshot frames (frame_count % 4 == 0), and the runs of frames between those are painted in
        // one parallel launch afterwards (recon_p_k). Ring output (a later snapshot frame would overwrite one still needed),
        // one-block-wide pictures (the last-block FILL reads the block's own previous pixels) and checksum calls walk
        // every block through every frame.
        steps.resize(F);
        rmeta.resize(F);
        rdst.resize(F);
        rstreams.resize(S);
        const uint32_t KA = cn / 4 + 2;                     // snapshot frames a stream can have in the chunk
        std::vector<RecMeta> ameta(par ? (size_t)KA * S : 0);
        std::vector<uint32_t*> adst(par ? (size_t)KA * S : 0, nullptr);
        std::vector<RecStream> astreams(par ? S : 0);
        std::vector<RecRun> runs;
        for (uint32_t s = 0; s < S; s++) {
            DecStream& d = ctx->streams[ids[s]];
            uint32_t na = 0;
            for (uint32_t k = 0; k < cn; k++) {
                const uint32_t g = d.next + k;       // global frame index == frame_count before this frame
                const uint32_t gi = g % 4 == 0 ? (g >= 4 ? g - 4 : EMPTY32) : g - g % 4;  // frame holding the I snapshot
                const uint32_t call_first = d.next - c0;  // first frame of this API call
                DecStep& x = steps[k * S + s];
                x.dst = outs[s] ? outs[s] + (size_t)(c0 + k) * P : d.d_ring + (size_t)(g % DEC_RING) * P;
                if (g == call_first) x.prev = d.d_img;
                else x.prev = outs[s] ? outs[s] + (size_t)(c0 + k - 1) * P : d.d_ring + (size_t)((g - 1) % DEC_RING) * P;
                if (gi == EMPTY32 || gi < call_first) x.ifr = d.d_ifr;
                else x.ifr = outs[s] ? outs[s] + (size_t)(gi - call_first) * P : d.d_ring + (size_t)(gi % DEC_RING) * P;
                x.fidx = s * cn + k;
                x.ebuf_off = fr[s * cn + k].ebuf_off;
                x.is_snap = g % 4 == 0;
                rmeta[k * S + s] = RecMeta{x.ebuf_off, x.fidx, k | (x.is_snap ? 0x80000000u : 0u)};
                rdst[k * S + s] = x.dst;
                if (!astreams.empty() && par) {
                    if (x.is_snap) {
                        if (na == 0) astreams[s] = RecStream{x.prev, x.ifr, d.d_pal, nullptr, d.dual, 0};
                        ameta[(size_t)na * S + s] = rmeta[k * S + s];
                        adst[(size_t)na * S + s] = x.dst;
                        na++;
                    } else if (!runs.empty() && runs.back().sidx == s && runs.back().first + runs.back().n == k && k > 0 && !steps[(k - 1) * S + s].is_snap) {
                        runs.back().n++;
                    } else {
                        runs.push_back(RecRun{x.ifr, x.prev, k, 1, s, 0});   // behind a snapshot frame of the chunk x.ifr is that frame's buffer
                    }
                }
            }
            if (runs.size() > 65535) par = false;   // (grid.y of recon_p_k) everything in order then: keeps is not passed on
            unsigned long long* ck = cks ? ctx->d_cksum.as<unsigned long long>() + (size_t)s * count + c0 : nullptr;
            rstreams[s] = RecStream{steps[s].prev, steps[s].ifr, d.d_pal, ck, d.dual, cn};
            if (!astreams.empty()) { astreams[s].cksum = ck; astreams[s].count = na; astreams[s].pal = d.d_pal; astreams[s].dual = d.dual; }
        }
        {
            // one upload: [all frames: position data, destinations, stream descriptors][snapshot frames: the same][runs]
            auto put = [&](const void* src, size_t bytes) {
                const size_t at = (stepbuf.size() + 15) & ~(size_t)15;
                stepbuf.resize(at + bytes);
                if (bytes) memcpy(stepbuf.data() + at, src, bytes);
                return at;
            };
            stepbuf.clear();
            const size_t o_meta = put(rmeta.data(), F * sizeof(RecMeta)), o_dst = put(rdst.data(), F * sizeof(uint32_t*)),
                         o_str = put(rstreams.data(), S * sizeof(RecStream));
            const size_t o_ameta = put(ameta.data(), ameta.size() * sizeof(RecMeta)), o_adst = put(adst.data(), adst.size() * sizeof(uint32_t*)),
                         o_astr = put(astreams.data(), astreams.size() * sizeof(RecStream)), o_runs = put(runs.data(), runs.size() * sizeof(RecRun));
            TRY(ensure(ctx, ctx->d_steps, stepbuf.size()));
            CK(cudaMemcpyAsync(ctx->d_steps.p, stepbuf.data(), stepbuf.size(), cudaMemcpyHostToDevice, ctx->st));
            const uint8_t* ds = ctx->d_steps.as<uint8_t>();
            const RecBase rb{ctx->d_recs.as<uint32_t>(), ctx->d_ebuf.as<uint8_t>(), ctx->d_bpos.as<uint32_t>(), ctx->d_stale.as<uint8_t>()};
            const uint8_t* keeps = par ? ctx->d_keeps.as<uint8_t>() : nullptr;
            const RecList all{reinterpret_cast<const RecMeta*>(ds + o_meta), reinterpret_cast<uint32_t* const*>(ds + o_dst), reinterpret_cast<const RecStream*>(ds + o_str)};
            const RecList snaps{reinterpret_cast<const RecMeta*>(ds + o_ameta), reinterpret_cast<uint32_t* const*>(ds + o_adst),
                                reinterpret_cast<const RecStream*>(ds + o_astr)};
            dim3 grid(cdiv(B, 128), S);
            static const int minb = getenv("AGMVB_RECON_MINB") ? atoi(getenv("AGMVB_RECON_MINB")) : 5;   // resident blocks per SM the kernel is compiled for
            if (minb >= 7) KL(ctx->lc, KC_RECON, (reconstruct_k<7><<<grid, 128, 0, ctx->st>>>(par ? snaps : all, all, rb, S, W, H, keeps)));
            else if (minb == 6) KL(ctx->lc, KC_RECON, (reconstruct_k<6><<<grid, 128, 0, ctx->st>>>(par ? snaps : all, all, rb, S, W, H, keeps)));
            else KL(ctx->lc, KC_RECON, (reconstruct_k<5><<<grid, 128, 0, ctx->st>>>(par ? snaps : all, all, rb, S, W, H, keeps)));
            if (par && !runs.empty()) {
                dim3 pgrid(cdiv(B, 128), (unsigned)runs.size());
                KL(ctx->lc, KC_RECON, (recon_p_k<<<pgrid, 128, 0, ctx->st>>>(reinterpret_cast<const RecRun*>(ds + o_runs), all.meta, all.dsts, all.streams, rb, S, W, H, keeps)));
            }
        }
        TRY(check_launch(ctx, "reconstruct"));
        // carry the state over: expanded-bitstream leftovers, last pixels, last I-frame snapshot
        for (uint32_t s = 0; s < S; s++) {
            DecStream& d = ctx->streams[ids[s]];
            // no frame of the chunk wrote at or beyond its expansion bound: only indices below the largest one can change
            uint64_t reach = 0;
            for (uint32_t k = 0; k < cn; k++) {
                const uint32_t g = d.next + k;
                const uint64_t e = d.lz77 ? std::min<uint64_t>((uint64_t)(d.csize[d.at(g)] / 4 + 1) * 256, d.persist_len) : d.usize[d.at(g)];
                reach = std::max<uint64_t>(reach, e + DEC_SLACK);
            }
            const uint32_t plen = (uint32_t)std::min<uint64_t>(d.persist_len, reach);
            if (plen)
                KL(ctx->lc, KC_STALE, (persist_update_k<<<cdiv(plen, 256), 256, 0, ctx->st>>>(dfr, ctx->d_bpos.as<uint32_t>(), s * cn, cn, ctx->d_ebuf.as<uint8_t>(),
                                                                                 d.d_persist, plen)));
        }
        if (c0 + cn == count) {
            for (uint32_t s = 0; s < S; s++) {
                DecStream& d = ctx->streams[ids[s]];
                const uint32_t call_first = d.next - c0, last = d.next + cn - 1;
                CK(cudaMemcpyAsync(d.d_img, steps[(cn - 1) * S + s].dst, P * 4, cudaMemcpyDeviceToDevice, ctx->st));
                const uint32_t gi = last - last % 4;  // most recent frame with frame_count % 4 == 0
                if (gi >= call_first) {
                    const uint32_t* src = outs[s] ? outs[s] + (size_t)(gi - call_first) * P : d.d_ring + (size_t)(gi % DEC_RING) * P;
                    CK(cudaMemcpyAsync(d.d_ifr, src, P * 4, cudaMemcpyDeviceToDevice, ctx->st));
                }
                CK(cudaMemcpyAsync(&d.last_bpos, ctx->d_bpos.as<uint32_t>() + (s * cn + cn - 1), 4, cudaMemcpyDeviceToHost, ctx->st));
                CK(cudaMemcpyAsync(&d.last_consumed, ctx->d_consumed.as<uint32_t>() + (s * cn + cn - 1), 4, cudaMemcpyDeviceToHost, ctx->st));
            }
        }
        if (S == 1 && ctx->info_bpos) CK(cudaMemcpyAsync(ctx->info_bpos + c0, ctx->d_bpos.p, (size_t)cn * 4, cudaMemcpyDeviceToHost, ctx->st));
        if (S == 1 && ctx->info_consumed) CK(cudaMemcpyAsync(ctx->info_consumed + c0, ctx->d_consumed.p, (size_t)cn * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));  // host vectors are rebuilt for the next chunk
        for (uint32_t s = 0; s < S; s++) ctx->streams[ids[s]].next += cn;
    }
    if (cks) {
        CK(cudaMemcpyAsync(cks, ctx->d_cksum.p, (size_t)S * count * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    }
    return OK;
}

extern "C" int agmvb_dec_batch(agmvb_ctx* ctx, const int* streams, uint32_t n_streams, uint32_t count, uint32_t* const* outs, uint64_t* checksums) {
    if (!ctx || !streams) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    std::vector<uint32_t*> none(n_streams, nullptr);
    return dec_batch_impl(ctx, streams, n_streams, count, outs ? outs : none.data(), checksums);
}

extern "C" int agmvb_dec_frames(agmvb_ctx* ctx, int stream, uint32_t count, uint32_t* out, int on_device) {
    if (!ctx || !out) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stream < 0 || (size_t)stream >= ctx->streams.size() || !ctx->streams[stream].open) FAIL(ERR_ARG, "bad stream handle");
    if (on_device) { uint32_t* o = out; return dec_batch_impl(ctx, &stream, 1, count, &o, nullptr); }
    const size_t P = (size_t)ctx->streams[stream].w * ctx->streams[stream].h;
    // host destination: decode into device staging, at most ~4 GB at a time (the expansion kernel is a warp per frame:
    // its latency is paid once per chunk, so chunks should hold as many frames as is reasonable)
    const uint32_t chunk = (uint32_t)std::max<size_t>(4, ((4ull << 30) / (P * 4)) & ~3ull);
    for (uint32_t c0 = 0; c0 < count; c0 += chunk) {
        uint32_t cn = std::min(chunk, count - c0);
        TRY(ensure(ctx, ctx->d_out, (size_t)cn * P * 4));
        uint32_t* o = ctx->d_out.as<uint32_t>();
        TRY(dec_batch_impl(ctx, &stream, 1, cn, &o, nullptr));
        if (ctx->host_fmt == 0) {
            GatedXfer gate(ctx, 1, ctx->st);
            CK(copy_pieces(out + (size_t)c0 * P, o, (size_t)cn * P * 4, cudaMemcpyDeviceToHost, ctx->st));
            CK(gate.finish(ctx->st));
        } else {
            const uint64_t npx = (uint64_t)cn * P;
            TRY(ensure(ctx, ctx->raw24, npx * 3 + 16));
            KL(ctx->lc, KC_MISC, (pack_bgr24_k<<<(unsigned)cdiv(cdiv(npx, 4), 256), 256, 0, ctx->st>>>(o, npx, ctx->raw24.as<uint8_t>())));
            GatedXfer gate(ctx, 1, ctx->st);
            CK(copy_pieces(reinterpret_cast<uint8_t*>(out) + (size_t)c0 * P * 3, ctx->raw24.p, npx * 3, cudaMemcpyDeviceToHost, ctx->st));
            CK(gate.finish(ctx->st));
        }
        CK(cudaStreamSynchronize(ctx->st));
    }
    return OK;
}

// ---- per-chunk entry for the reference's streaming callers ------------------------
// AGMV_DecodeFrameChunk (src/agmv_decode.c:145-410) is called with a FILE* at 'AGFC' by AGMV_PlayAGMV
// (src/agmv_playback.c:102-115) and players; the drop-in reads the chunk and hands the payload here.
extern "C" int agmvb_dec_open_raw(agmvb_ctx* ctx, uint32_t w, uint32_t h, int version, const uint32_t pal0[256], const uint32_t pal1[256],
                                  int* stream) {
    if (!ctx || !stream || !pal0) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (version < 1 || version > 4) FAIL(ERR_HEADER, "bad stream version %d", version);
    if (w == 0 || h == 0 || (w & 3) || (h & 3)) FAIL(ERR_UNSUPPORTED, "width and height must be multiples of 4");
    DecStream s;
    s.w = w; s.h = h; s.n_frames = 0xFFFFFFFFu; s.version = (uint32_t)version; s.dual = version == 1 || version == 3; s.lz77 = version >= 3;
    s.raw = true;
    s.data_off.assign(1, 0); s.usize.assign(1, 0); s.csize.assign(1, 0);
    const size_t P = (size_t)w * h;
    uint32_t pal[512];
    memset(pal, 0, sizeof pal);
    for (int i = 0; i < 256; i++) pal[i] = pal0[i] & 0xFFFFFFu;
    if (s.dual && pal1) for (int i = 0; i < 256; i++) pal[256 + i] = pal1[i] & 0xFFFFFFu;
    CK(cudaMalloc(&s.d_pal, 512 * 4));
    CK(cudaMemcpyAsync(s.d_pal, pal, 512 * 4, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMalloc(&s.d_img, P * 4));
    CK(cudaMalloc(&s.d_ifr, P * 4));
    s.persist_len = (uint32_t)(2 * P + 64);
    CK(cudaMalloc(&s.d_persist, s.persist_len));
    CK(cudaMemsetAsync(s.d_img, 0, P * 4, ctx->st));
    CK(cudaMemsetAsync(s.d_ifr, 0, P * 4, ctx->st));
    CK(cudaMemsetAsync(s.d_persist, 0, s.persist_len, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    s.open = true;
    int id = -1;
    for (size_t k = 0; k < ctx->streams.size(); k++) if (!ctx->streams[k].open) { id = (int)k; break; }
    if (id < 0) { ctx->streams.push_back(DecStream()); id = (int)ctx->streams.size() - 1; }
    ctx->streams[id] = s;
    *stream = id;
    return OK;
}

// payload: the bytes that follow the 16-byte chunk header in the file (csize bytes plus whatever comes after,
// at least 8 more if available: the bit reader may run into the trailer). frame_count: agmv->frame_count before
// this frame. Outputs: pixels (host, w*h), bitstream->pos, and how many payload bytes the reader consumed.
extern "C" int agmvb_dec_chunk(agmvb_ctx* ctx, int stream, const uint8_t* payload, uint64_t payload_len, uint32_t usize, uint32_t csize,
                               uint32_t frame_count, uint32_t* out_px, uint32_t* bpos, uint32_t* consumed) {
    if (!ctx || !payload || !out_px) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stream < 0 || (size_t)stream >= ctx->streams.size() || !ctx->streams[stream].open || !ctx->streams[stream].raw)
        FAIL(ERR_ARG, "bad raw stream handle");
    DecStream& d = ctx->streams[stream];
    const size_t P = (size_t)d.w * d.h;
    if (usize > 2 * P + 64) FAIL(ERR_MEMORY, "uncompressed size %u exceeds the reference's bitstream buffer", usize);
    if (payload_len + 64 > d.file_cap) {
        CK(cudaStreamSynchronize(ctx->st));
        if (d.d_file) CK(cudaFree(d.d_file));
        d.file_cap = payload_len * 2 + 4096;
        CK(cudaMalloc(&d.d_file, d.file_cap));
    }
    CK(cudaMemcpyAsync(d.d_file, payload, payload_len, cudaMemcpyHostToDevice, ctx->st));
    d.file_len = payload_len;
    d.data_off.assign(1, 0); d.usize.assign(1, usize); d.csize.assign(1, csize);
    d.next = d.raw_base = frame_count;
    TRY(ensure(ctx, ctx->d_out, P * 4));
    uint32_t* o = ctx->d_out.as<uint32_t>();
    TRY(dec_batch_impl(ctx, &stream, 1, 1, &o, nullptr));
    CK(cudaMemcpyAsync(out_px, o, P * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (bpos) *bpos = d.last_bpos;
    if (consumed) *consumed = d.last_consumed;
    return OK;
}

// n consecutive frame chunks of one stream in one call - the frame-ahead queue behind AGMV_PlayAGMV's loop
// (src/agmv_playback.c:102-115). slab: a piece of the file that holds all n chunks, payload_off[k] = offset in
// the slab of the first byte after chunk k's 16-byte header (so the bit reader sees the real bytes after a payload, as
// fread does). Outputs per frame: pixels (host, n*w*h), bitstream->pos, payload bytes consumed.
extern "C" int agmvb_dec_chunks(agmvb_ctx* ctx, int stream, const uint8_t* slab, uint64_t slab_len, uint32_t n, const uint64_t* payload_off,
                                const uint32_t* usize, const uint32_t* csize, uint32_t first_frame_count, uint32_t* out_px, uint32_t* bpos,
                                uint32_t* consumed) {
    if (!ctx || !slab || !payload_off || !usize || !csize || !out_px || n == 0) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stream < 0 || (size_t)stream >= ctx->streams.size() || !ctx->streams[stream].open || !ctx->streams[stream].raw)
        FAIL(ERR_ARG, "bad raw stream handle");
    DecStream& d = ctx->streams[stream];
    const size_t P = (size_t)d.w * d.h;
    for (uint32_t k = 0; k < n; k++) {
        if (usize[k] > 2 * P + 64) FAIL(ERR_MEMORY, "uncompressed size %u exceeds the reference's bitstream buffer", usize[k]);
        if (payload_off[k] > slab_len) FAIL(ERR_ARG, "chunk %u lies outside the slab", k);
    }
    if (slab_len + 64 > d.file_cap) {
        CK(cudaStreamSynchronize(ctx->st));
        if (d.d_file) CK(cudaFree(d.d_file));
        d.file_cap = slab_len * 2 + 4096;
        CK(cudaMalloc(&d.d_file, d.file_cap));
    }
    CK(cudaMemcpyAsync(d.d_file, slab, slab_len, cudaMemcpyHostToDevice, ctx->st));
    d.file_len = slab_len;
    d.data_off.assign(payload_off, payload_off + n); d.usize.assign(usize, usize + n); d.csize.assign(csize, csize + n);
    d.next = d.raw_base = first_frame_count;
    TRY(ensure(ctx, ctx->d_out, (size_t)n * P * 4));
    uint32_t* o = ctx->d_out.as<uint32_t>();
    ctx->info_bpos = bpos; ctx->info_consumed = consumed;
    const int rc = dec_batch_impl(ctx, &stream, 1, n, &o, nullptr);
    ctx->info_bpos = ctx->info_consumed = nullptr;
    if (rc) return rc;
    CK(cudaMemcpyAsync(out_px, o, (size_t)n * P * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

// Decoder state of a stream (what the reference keeps in the AGMV handle: frame pixels, I-frame copy, bitstream buffer) put
// aside / brought back, so that frames decoded ahead of the caller can be withdrawn when the caller seeks instead.
extern "C" int agmvb_dec_snapshot(agmvb_ctx* ctx, int stream) {
    if (!ctx) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stream < 0 || (size_t)stream >= ctx->streams.size() || !ctx->streams[stream].open) FAIL(ERR_ARG, "bad stream handle");
    DecStream& d = ctx->streams[stream];
    const size_t P = (size_t)d.w * d.h;
    if (!d.d_snap) CK(cudaMalloc(&d.d_snap, P * 8 + d.persist_len));
    CK(cudaMemcpyAsync(d.d_snap, d.d_img, P * 4, cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaMemcpyAsync(d.d_snap + P * 4, d.d_ifr, P * 4, cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaMemcpyAsync(d.d_snap + P * 8, d.d_persist, d.persist_len, cudaMemcpyDeviceToDevice, ctx->st));
    d.snap_valid = true;
    return OK;
}
extern "C" int agmvb_dec_restore(agmvb_ctx* ctx, int stream) {
    if (!ctx) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (stream < 0 || (size_t)stream >= ctx->streams.size() || !ctx->streams[stream].open || !ctx->streams[stream].snap_valid)
        FAIL(ERR_ARG, "no snapshot to restore");
    DecStream& d = ctx->streams[stream];
    const size_t P = (size_t)d.w * d.h;
    CK(cudaMemcpyAsync(d.d_img, d.d_snap, P * 4, cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaMemcpyAsync(d.d_ifr, d.d_snap + P * 4, P * 4, cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaMemcpyAsync(d.d_persist, d.d_snap + P * 8, d.persist_len, cudaMemcpyDeviceToDevice, ctx->st));
    return OK;
}

// ===========================================================================
// bench / test utilities
// ===========================================================================
// Deterministic synthetic frames (SURVEY.md 8d / BASELINE.md 4), generated on the device so that large
// benchmark inputs need not be produced on the host. Same integer formula as oracle/agmv_oracle.c:orc_synth_frame.
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 13;
    h *= 0x5bd1e995u;
    h ^= h >> 15;
    return h;
}
__global__ void synth_k(uint32_t* __restrict__ out, int w, int h, int first_t, uint32_t seed) {
    const int t = first_t + blockIdx.y;
    const size_t P = (size_t)w * h;
    const int s = w / 8 > 8 ? w / 8 : 8;
    const int mx = w - s > 1 ? w - s : 1, my = h - s > 1 ? h - s : 1;
    const int sx = (5 * t) % mx, sy = (3 * t) % my;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(p % w), y = (int)(p / w);
        uint32_t r, g, b;
        if (x >= sx && x < sx + s && y >= sy && y < sy + s) {
            uint32_t nh = mix32(((uint32_t)x * 73856093u) ^ ((uint32_t)y * 19349663u) ^ ((uint32_t)t * 83492791u) ^ seed);
            r = nh & 255; g = (nh >> 8) & 255; b = (nh >> 16) & 255;
        } else if (y >= h / 3 && y < 2 * h / 3) {
            r = (uint32_t)(x * 255 / (w - 1) + 2 * t) & 255;
            g = (uint32_t)(y * 255 / (h - 1)) & 255;
            b = (uint32_t)(x + y + 4 * t) & 255;
        } else {
            uint32_t th = mix32(((uint32_t)(x / 32) * 73856093u) ^ ((uint32_t)(y / 32) * 19349663u) ^ seed);
            r = th & 255; g = (th >> 8) & 255; b = (th >> 16) & 255;
        }
        out[(size_t)blockIdx.y * P + p] = r << 16 | g << 8 | b;
    }
}

extern "C" int agmvb_synth_frames(agmvb_ctx* ctx, uint32_t* dev_out, uint32_t w, uint32_t h, uint32_t first_t, uint32_t n, uint32_t seed) {
    if (!ctx || !dev_out || !w || !h) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    for (uint32_t f0 = 0; f0 < n; f0 += 32768) {
        uint32_t nf = std::min<uint32_t>(32768, n - f0);
        dim3 grid(std::min<uint32_t>(cdiv((size_t)w * h, 256), 1184), nf);
        KL(ctx->lc, KC_MISC, (synth_k<<<grid, 256, 0, ctx->st>>>(dev_out + (size_t)f0 * w * h, (int)w, (int)h, (int)(first_t + f0), seed)));
    }
    return check_launch(ctx, "synth");
}

// Per-kernel-class device time from CUDA events recorded around every launch on the launching stream.
extern "C" int agmvb_profile(agmvb_ctx* ctx, int enable) {
    if (!ctx) return ERR_ARG;
    CK(cudaStreamSynchronize(ctx->st));
    ctx->lc.prof = enable != 0;
    ctx->lc.nrec = 0;
    ctx->lc.npool = 0;
    return OK;
}
extern "C" int agmvb_profile_classes(void) { return KC_COUNT; }
extern "C" const char* agmvb_profile_name(int cls) { return kclass_name(cls); }
// Sums and clears what was recorded since the last read: per class, launch count and total milliseconds.
extern "C" int agmvb_profile_read(agmvb_ctx* ctx, uint64_t* counts, double* total_ms) {
    if (!ctx || !counts || !total_ms) return ERR_ARG;
    CK(cudaStreamSynchronize(ctx->st));
    for (int c = 0; c < KC_COUNT; c++) { counts[c] = 0; total_ms[c] = 0.0; }
    for (size_t k = 0; k < ctx->lc.nrec; k++) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->lc.recs[k].a, ctx->lc.recs[k].b));
        counts[ctx->lc.recs[k].cls]++;
        total_ms[ctx->lc.recs[k].cls] += ms;
    }
    ctx->lc.nrec = 0;
    ctx->lc.npool = 0;
    return OK;
}

// ===========================================================================
// unit-test hooks
// ===========================================================================
// Copy an internal decode buffer of the last decoded chunk to the host (debugging / white-box tests):
// 0 = bpos per frame, 1 = block records, 2 = consumed bytes per frame, 3 = stale bytes, 4 = expanded bitstreams.
extern "C" int agmvb_test_peek(agmvb_ctx* ctx, int which, void* dst, uint64_t bytes) {
    if (!ctx || !dst) return ERR_ARG;
    DBuf* b = which == 0 ? &ctx->d_bpos : which == 1 ? &ctx->d_recs : which == 2 ? &ctx->d_consumed : which == 3 ? &ctx->d_stale : &ctx->d_ebuf;
    if (!b->p || bytes > b->cap) FAIL(ERR_ARG, "peek: buffer %d has %zu bytes", which, b->cap);
    CK(cudaMemcpyAsync(dst, b->p, bytes, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

extern "C" int agmvb_test_lzss(agmvb_ctx* ctx, const uint8_t* data, const uint32_t* frame_start, uint32_t F, uint8_t* out, uint64_t out_cap,
                               uint64_t* out_off, uint32_t* csize, uint32_t* outbits) {
    if (!ctx || !frame_start || F == 0) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint32_t n = frame_start[F];
    TRY(ensure(ctx, ctx->bs, (size_t)n + 256));
    CK(cudaMemsetAsync(ctx->bs.p, 0, (size_t)n + 256, ctx->st));
    if (n) CK(cudaMemcpyAsync(ctx->bs.p, data, n, cudaMemcpyHostToDevice, ctx->st));
    ctx->image_bytes = 0;
    ctx->last_usize.clear();
    ctx->last_csize.clear();
    TRY(lz_group(ctx, ctx->bs.as<uint8_t>(), frame_start, F, 0));
    std::vector<uint32_t> ob(F), wb(F + 1);
    CK(cudaMemcpyAsync(ob.data(), ctx->lz.outbits, (size_t)F * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(wb.data(), ctx->lz.wbase, (size_t)(F + 1) * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    uint64_t o = 0;
    for (uint32_t f = 0; f < F; f++) {
        uint64_t nbytes = ((uint64_t)ob[f] + 7) / 8;
        if (o + nbytes > out_cap) FAIL(ERR_ARG, "output buffer too small");
        if (nbytes) CK(cudaMemcpyAsync(out + o, ctx->lz.out_words + wb[f], nbytes, cudaMemcpyDeviceToHost, ctx->st));
        if (out_off) out_off[f] = o;
        if (csize) csize[f] = ctx->last_csize[f];
        if (outbits) outbits[f] = ob[f];
        o += nbytes;
    }
    if (out_off) out_off[F] = o;
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

// AGMV_LZ77 (src/agmv_encode.c:179-238) over F buffers sharing one carried bitstream buffer whose bytes all start as `persist_fill`
extern "C" int agmvb_test_lz77(agmvb_ctx* ctx, const uint8_t* data, const uint32_t* frame_start, uint32_t F, int persist_fill, uint8_t* out,
                               uint64_t out_cap, uint64_t* out_off, uint32_t* csize) {
    if (!ctx || !frame_start || F == 0) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint32_t n = frame_start[F];
    uint32_t longest = 0;
    for (uint32_t f = 0; f < F; f++) longest = std::max(longest, frame_start[f + 1] - frame_start[f]);
    TRY(ensure(ctx, ctx->bs, (size_t)n + 256));
    CK(cudaMemsetAsync(ctx->bs.p, 0, (size_t)n + 256, ctx->st));
    if (n) CK(cudaMemcpyAsync(ctx->bs.p, data, n, cudaMemcpyHostToDevice, ctx->st));
    TRY(ensure(ctx, ctx->l77_persist, (size_t)longest + 64));
    CK(cudaMemsetAsync(ctx->l77_persist.p, persist_fill, (size_t)longest + 64, ctx->st));
    ctx->image_bytes = 0;
    ctx->last_usize.clear();
    ctx->last_csize.clear();
    TRY(lz77_group(ctx, ctx->bs.as<uint8_t>(), frame_start, F, 0));
    const size_t stride = (size_t)F + 2;
    std::vector<uint32_t> ob(F), wb(F);
    CK(cudaMemcpyAsync(wb.data(), ctx->l77_meta.as<uint32_t>() + stride, (size_t)F * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(ob.data(), ctx->l77_meta.as<uint32_t>() + stride * 4, (size_t)F * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    uint64_t o = 0;
    for (uint32_t f = 0; f < F; f++) {
        uint64_t nbytes = ob[f] / 8;
        if (o + nbytes > out_cap) FAIL(ERR_ARG, "output buffer too small");
        if (nbytes) CK(cudaMemcpyAsync(out + o, ctx->l77_out.as<uint32_t>() + wb[f], nbytes, cudaMemcpyDeviceToHost, ctx->st));
        if (out_off) out_off[f] = o;
        if (csize) csize[f] = ctx->last_csize[f];
        o += nbytes;
    }
    if (out_off) out_off[F] = o;
    CK(cudaStreamSynchronize(ctx->st));
    return OK;
}

extern "C" int agmvb_test_quantize(agmvb_ctx* ctx, const uint32_t* colors, uint64_t n, const uint32_t pal0[256], const uint32_t pal1[256],
                                   int dual, uint16_t* entries) {
    if (!ctx || !colors || !pal0 || !entries || (n & 3) || n > 0xFFFFFFFCull) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const bool fill = dual & 2;   // bit 1: compute the whole table first (lut_fill_k), as long sequences do
    dual &= 1;
    uint32_t pal[512];
    memset(pal, 0, sizeof pal);
    memcpy(pal, pal0, 1024);
    if (dual && pal1) memcpy(pal + 256, pal1, 1024);
    for (int i = 0; i < 512; i++) pal[i] &= 0xFFFFFFu;
    uint32_t* d_pal; uint16_t* d_lut;
    CK(cudaMalloc(&d_pal, 2048));
    CK(cudaMalloc(&d_lut, sizeof(uint16_t) << 24));
    CK(cudaMemsetAsync(d_lut, 0xFF, sizeof(uint16_t) << 24, ctx->st));
    CK(cudaMemcpyAsync(d_pal, pal, 2048, cudaMemcpyHostToDevice, ctx->st));
    TRY(ensure(ctx, ctx->stage, n * 4));
    TRY(ensure(ctx, ctx->entries, n * 2));
    TRY(ensure(ctx, ctx->srcpairs, sizeof(SrcPair)));
    CK(cudaMemcpyAsync(ctx->stage.p, colors, n * 4, cudaMemcpyHostToDevice, ctx->st));
    SrcPair sp{ctx->stage.as<uint32_t>(), nullptr};
    CK(cudaMemcpyAsync(ctx->srcpairs.p, &sp, sizeof sp, cudaMemcpyHostToDevice, ctx->st));
    dim3 grid(cdiv(n / 4, 256), 1);
    if (fill) KL(ctx->lc, KC_QUANT, (lut_fill_k<<<(1u << 24) / 1024u, 256, 0, ctx->st>>>(d_pal, dual ? 512 : 256, reinterpret_cast<uint2*>(d_lut))));
    KL(ctx->lc, KC_QUANT, (quantize_k<<<grid, 256, 0, ctx->st>>>(ctx->srcpairs.as<SrcPair>(), nullptr, (uint32_t)n, d_pal, dual ? 512 : 256, d_lut, ctx->entries.as<uint16_t>())));
    TRY(check_launch(ctx, "quantize"));
    CK(cudaMemcpyAsync(entries, ctx->entries.p, n * 2, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    cudaFree(d_pal);
    cudaFree(d_lut);
    return OK;
}

extern "C" int agmvb_test_assemble(agmvb_ctx* ctx, const uint16_t* entries, const uint16_t* iframe_entries, uint32_t w, uint32_t h, int dual,
                                   const uint32_t pal0[256], const uint32_t pal1[256], uint8_t* out, uint64_t cap, uint32_t* usize) {
    if (!ctx || !entries || !pal0 || !out || (w & 3) || (h & 3)) return ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t P = (size_t)w * h;
    const uint32_t B = (w >> 2) * (h >> 2);
    uint32_t pal[512];
    memset(pal, 0, sizeof pal);
    memcpy(pal, pal0, 1024);
    if (pal1) memcpy(pal + 256, pal1, 1024);
    uint32_t* d_pal;
    CK(cudaMalloc(&d_pal, 2048));
    CK(cudaMemcpyAsync(d_pal, pal, 2048, cudaMemcpyHostToDevice, ctx->st));
    TRY(ensure(ctx, ctx->entries, P * 4));
    uint16_t* d_e = ctx->entries.as<uint16_t>();
    CK(cudaMemcpyAsync(d_e, entries, P * 2, cudaMemcpyHostToDevice, ctx->st));
    if (iframe_entries) CK(cudaMemcpyAsync(d_e + P, iframe_entries, P * 2, cudaMemcpyHostToDevice, ctx->st));
    EntPair ep{d_e, iframe_entries ? d_e + P : nullptr};
    TRY(ensure(ctx, ctx->entpairs, sizeof ep));
    TRY(ensure(ctx, ctx->rec, B));
    TRY(ensure(ctx, ctx->boff, (size_t)B * 4));
    TRY(ensure(ctx, ctx->scanws, ((size_t)cdiv(B, SCAN_TILE) + 2) * 4));
    TRY(ensure(ctx, ctx->bs, (size_t)B * 33 + 256));
    TRY(ensure(ctx, ctx->fs, 16));
    CK(cudaMemcpyAsync(ctx->entpairs.p, &ep, sizeof ep, cudaMemcpyHostToDevice, ctx->st));
    dim3 grid(cdiv(B, 256), 1);
    KL(ctx->lc, KC_CLASSIFY, (classify_k<<<grid, 256, 0, ctx->st>>>(ctx->entpairs.as<EntPair>(), w, h, d_pal, dual, ctx->rec.as<uint8_t>())));
    device_scan<SumOp, true>(RecLen{ctx->rec.as<uint8_t>()}, StoreU32{ctx->boff.as<uint32_t>()}, B, ctx->scanws.as<uint32_t>(), ctx->lc, KC_BLOCKSCAN);
    KL(ctx->lc, KC_BLOCKSCAN, (frame_starts_k<<<1, 256, 0, ctx->st>>>(ctx->boff.as<uint32_t>(), B, 1, ctx->scanws.as<uint32_t>() + cdiv(B, SCAN_TILE), ctx->fs.as<uint32_t>())));
    KL(ctx->lc, KC_EMIT, (emit_k<<<grid, 256, 0, ctx->st>>>(ctx->entpairs.as<EntPair>(), w, h, dual, ctx->rec.as<uint8_t>(), ctx->boff.as<uint32_t>(), ctx->bs.as<uint8_t>())));
    TRY(check_launch(ctx, "assemble"));
    uint32_t fs[2];
    CK(cudaMemcpyAsync(fs, ctx->fs.p, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (fs[1] > cap) { cudaFree(d_pal); FAIL(ERR_ARG, "output buffer too small"); }
    CK(cudaMemcpyAsync(out, ctx->bs.p, fs[1], cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (usize) *usize = fs[1];
    cudaFree(d_pal);
    return OK;
}
