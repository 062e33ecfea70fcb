// K4: exact LZSS for a batch of frame bitstreams (reference: AGMV_LZSS,
// src/agmv_encode.c:106-177, bit writer src/agmv_utils.c:86-112, chunk framing
// src/agmv_encode.c:549-624).
//
// The reference does a brute-force search per parse position: longest match
// (3..15 bytes, overlap allowed) inside the previous 65535 bytes, and among
// equal lengths the EARLIEST start. That is O(n * 65535) per frame. Here the
// same answer comes from 15 stable counting-sort passes ("radix refinement"):
//
//   level L array A_L = all positions of the batch grouped by their first L
//   bytes, ascending position inside a group. A_{L+1} is one stable 8-bit pass
//   of A_L keyed on byte[p+L]; equal (old group, digit) pairs stay contiguous,
//   so group starts are a running max over head flags.
//   * a match of length >= L exists for position p  <=>  p's predecessor in its
//     level-L group is within the window (the nearest previous occurrence),
//     so bestlen[p] = max such L (monotone in L, written level by level).
//   * the earliest start is a lower-bound search for p-65535 inside the group
//     (positions are sorted), done only for positions the greedy parse visits.
//
// The greedy parse itself is a pointer chase (i += len or 1). It is resolved
// tile-parallel: every 1024-position tile is walked speculatively from each of
// the 15 possible entry offsets, one CTA then chains the tiles, and a last pass
// re-walks each tile from its real entry writing the running bit cursor.
// Bits are packed LSB-first with atomicOr on 32-bit words; csize uses the same
// float expression as the reference; the 8x0xFF trailer is written at
// data+csize so it clobbers the last partial byte exactly like the reference.
#pragma once
#include "common.cuh"
#include "radix.cuh"
#include "scan.cuh"

namespace agmvb {

constexpr int LZ_LEVELS = 15;
constexpr int PARSE_TILE = 1024;
constexpr int PARSE_CHAIN_CHUNK = 512;

struct LzWork {
    uint32_t cap_n = 0, cap_frames = 0;
    uint8_t* maxlen = nullptr;
    uint8_t* bestlen = nullptr;
    uint32_t* A[LZ_LEVELS + 1] = {};
    uint32_t* gs[2] = {};
    uint32_t* lvlidx = nullptr;
    uint32_t* gsat = nullptr;
    uint32_t* bitcum = nullptr;   // cap_n + 1
    uint32_t* tile_hist = nullptr;
    uint32_t* scan_ws = nullptr;
    uint8_t* exit_tab = nullptr;  // ntile * 16
    uint16_t* w_tab = nullptr;    // ntile * 16
    uint8_t* entry_tab = nullptr; // ntile
    uint32_t* cumbase = nullptr;  // ntile
    uint32_t* out_words = nullptr;
    size_t out_words_cap = 0;
    uint32_t* wbase = nullptr;    // cap_frames + 1
    uint32_t* outbits = nullptr;  // cap_frames
    uint32_t* csize = nullptr;    // cap_frames
    uint32_t* chunk_off = nullptr;// cap_frames + 1
};

struct APtrs { const uint32_t* a[LZ_LEVELS + 1]; };

__device__ __forceinline__ uint32_t frame_of(const uint32_t* __restrict__ fs, uint32_t F, uint32_t i) {
    // largest f with fs[f] <= i and fs[f+1] > i (empty frames are skipped)
    uint32_t lo = 0, hi = F;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (fs[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void lz_init_k(uint32_t n, const uint32_t* __restrict__ fs, uint32_t F, uint32_t* __restrict__ A0,
                          uint32_t* __restrict__ gs0, uint8_t* __restrict__ maxlen, uint8_t* __restrict__ bestlen,
                          uint32_t* __restrict__ bitcum, uint32_t* __restrict__ wbase) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= F) wbase[i] = (uint32_t)(((uint64_t)fs[i] * 9u) >> 5) + 3u * i;
    if (i == n) bitcum[n] = EMPTY32;
    if (i >= n) return;
    uint32_t f = frame_of(fs, F, i);
    uint32_t rem = fs[f + 1] - i;
    A0[i] = i;
    gs0[i] = fs[f];
    maxlen[i] = (uint8_t)(rem < (uint32_t)LZ_MAXLEN ? rem : (uint32_t)LZ_MAXLEN);
    bestlen[i] = 0;
    bitcum[i] = EMPTY32;
}

struct LzDigit {
    const uint8_t* bs;
    const uint32_t* pos;
    uint32_t L;
    __device__ uint32_t operator()(uint32_t i) const { return bs[pos[i] + L]; }
};
struct LzMove {
    const uint32_t* pos_in;
    const uint32_t* gs_in;
    uint32_t* pos_out;
    uint32_t* gs_out;
    __device__ void operator()(uint32_t s, uint32_t d) const {
        pos_out[d] = pos_in[s];
        gs_out[d] = gs_in[s];
    }
};
// head flag of the refined array: value idx at group heads, 0 elsewhere
struct LzHead {
    const uint8_t* bs;
    const uint32_t* pos;     // level L+1 order
    const uint32_t* gs_old;  // level L group starts carried through the scatter
    uint32_t L;
    __device__ uint32_t operator()(uint32_t idx) const {
        if (idx == 0) return 0;
        bool head = gs_old[idx] != gs_old[idx - 1] || bs[pos[idx] + L] != bs[pos[idx - 1] + L];
        return head ? idx : 0u;
    }
};
struct LzGroupOut {
    const uint32_t* pos;
    uint32_t* gs_new;
    const uint8_t* maxlen;
    uint8_t* bestlen;
    uint32_t* lvlidx;
    uint32_t* gsat;
    uint32_t Lnew;
    __device__ void operator()(uint32_t idx, uint32_t g) const {
        gs_new[idx] = g;
        if (Lnew >= (uint32_t)LZ_MINLEN && g != idx) {
            uint32_t p = pos[idx], prev = pos[idx - 1];
            if (p - prev <= (uint32_t)LZ_WINDOW && Lnew <= maxlen[p]) {
                bestlen[p] = (uint8_t)Lnew;
                lvlidx[p] = idx;
                gsat[p] = g;
            }
        }
    }
};

// ---- greedy parse, tile-parallel ------------------------------------------
__global__ void __launch_bounds__(256) lz_parse_spec_k(const uint8_t* __restrict__ bestlen, uint32_t n, uint32_t ntile,
                                                       uint8_t* __restrict__ exit_tab, uint16_t* __restrict__ w_tab) {
    __shared__ __align__(16) uint8_t s[8][PARSE_TILE];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t tile = blockIdx.x * 8 + warp;
    if (tile >= ntile) return;
    const uint32_t t0 = tile * PARSE_TILE;
    const uint32_t end = min(t0 + PARSE_TILE, n);
    for (uint32_t k = lane; k < PARSE_TILE; k += 32) s[warp][k] = (t0 + k < n) ? bestlen[t0 + k] : 0;
    __syncwarp();
    if (lane < LZ_MAXLEN) {
        uint32_t i = t0 + lane, w = 0;
        while (i < end) {
            uint32_t l = s[warp][i - t0];
            if (l >= (uint32_t)LZ_MINLEN) { w += 21; i += l; } else { w += 9; i += 1; }
        }
        exit_tab[tile * 16 + lane] = (uint8_t)(i - end);
        w_tab[tile * 16 + lane] = (uint16_t)w;
    }
}

__global__ void __launch_bounds__(1024) lz_parse_chain_k(uint32_t ntile, const uint8_t* __restrict__ exit_tab,
                                                         const uint16_t* __restrict__ w_tab, uint8_t* __restrict__ entry_tab,
                                                         uint32_t* __restrict__ cumbase) {
    __shared__ uint8_t se[PARSE_CHAIN_CHUNK * 16];
    __shared__ uint16_t sw[PARSE_CHAIN_CHUNK * 16];
    __shared__ uint32_t s_e, s_cum;
    if (threadIdx.x == 0) { s_e = 0; s_cum = 0; }
    for (uint32_t c0 = 0; c0 < ntile; c0 += PARSE_CHAIN_CHUNK) {
        uint32_t cnt = min((uint32_t)PARSE_CHAIN_CHUNK, ntile - c0);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < cnt * 16; k += 1024) {
            se[k] = exit_tab[(size_t)c0 * 16 + k];
            sw[k] = w_tab[(size_t)c0 * 16 + k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t e = s_e, cum = s_cum;
            for (uint32_t k = 0; k < cnt; k++) {
                entry_tab[c0 + k] = (uint8_t)e;
                cumbase[c0 + k] = cum;
                cum += sw[k * 16 + e];
                e = se[k * 16 + e];
            }
            s_e = e; s_cum = cum;
        }
    }
}

__global__ void __launch_bounds__(256) lz_parse_mark_k(const uint8_t* __restrict__ bestlen, uint32_t n, uint32_t ntile,
                                                       const uint8_t* __restrict__ entry_tab, const uint32_t* __restrict__ cumbase,
                                                       uint32_t* __restrict__ bitcum) {
    __shared__ __align__(16) uint8_t s[8][PARSE_TILE];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t tile = blockIdx.x * 8 + warp;
    if (tile >= ntile) return;
    const uint32_t t0 = tile * PARSE_TILE;
    const uint32_t end = min(t0 + PARSE_TILE, n);
    for (uint32_t k = lane; k < PARSE_TILE; k += 32) s[warp][k] = (t0 + k < n) ? bestlen[t0 + k] : 0;
    __syncwarp();
    if (lane == 0) {
        uint32_t i = t0 + entry_tab[tile], cum = cumbase[tile];
        while (i < end) {
            bitcum[i] = cum;
            uint32_t l = s[warp][i - t0];
            if (l >= (uint32_t)LZ_MINLEN) { cum += 21; i += l; } else { cum += 9; i += 1; }
        }
        if (end == n) bitcum[n] = cum;  // the orbit ends exactly at n
    }
}

// ---- token emission ----------------------------------------------------------
__global__ void __launch_bounds__(256) lz_pack_k(const uint8_t* __restrict__ bs, uint32_t n, const uint32_t* __restrict__ fs, uint32_t F,
                                                 const uint8_t* __restrict__ bestlen, const uint32_t* __restrict__ lvlidx,
                                                 const uint32_t* __restrict__ gsat, const uint32_t* __restrict__ bitcum, APtrs A,
                                                 const uint32_t* __restrict__ wbase, uint32_t* __restrict__ out_words) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = bitcum[i];
    if (c == EMPTY32) return;
    uint32_t f = frame_of(fs, F, i);
    uint32_t rel = c - bitcum[fs[f]];
    uint32_t l = bestlen[i], v, nb;
    if (l >= (uint32_t)LZ_MINLEN) {
        const uint32_t* a = A.a[l];
        uint32_t lo = gsat[i], hi = lvlidx[i];
        uint32_t target = i > (uint32_t)LZ_WINDOW ? i - (uint32_t)LZ_WINDOW : 0u;
        while (lo < hi) {  // first j in [lo,hi) with a[j] >= target; a[hi-1] qualifies by construction
            uint32_t mid = (lo + hi) >> 1;
            if (a[mid] >= target) hi = mid; else lo = mid + 1;
        }
        uint32_t off = i - a[lo];
        v = (off << 1) | (l << 17);
        nb = 21;
    } else {
        v = 1u | ((uint32_t)bs[i] << 1);
        nb = 9;
    }
    uint32_t w = wbase[f] + (rel >> 5), sh = rel & 31;
    atomicOr(&out_words[w], v << sh);
    if (sh + nb > 32) atomicOr(&out_words[w + 1], v >> (32 - sh));
}

__global__ void __launch_bounds__(1024) lz_finalize_k(const uint32_t* __restrict__ fs, uint32_t F, const uint32_t* __restrict__ bitcum,
                                                      uint32_t* __restrict__ outbits, uint32_t* __restrict__ csize,
                                                      uint32_t* __restrict__ chunk_off) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < F; base += 1024) {
        uint32_t f = base + threadIdx.x;
        uint32_t len = 0;
        if (f < F) {
            uint32_t a = fs[f], b = fs[f + 1];
            uint32_t ob = (a == b) ? 0u : bitcum[b] - bitcum[a];
            uint32_t cs = (uint32_t)((float)(int)ob / 8.0f);  // src/agmv_encode.c:176
            outbits[f] = ob;
            csize[f] = cs;
            len = 32u + cs;  // AGFC hdr 16 + payload + 8 x 0xFF + empty AGAC chunk 8
        }
        uint32_t inc = warp_inclusive<SumOp>(len);
        if (lane_id() == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) wsum[threadIdx.x] = warp_inclusive<SumOp>(wsum[threadIdx.x]);
        __syncthreads();
        uint32_t pre = carry_s + ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0u);
        if (f < F) chunk_off[f] = pre + inc - len;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = pre + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_off[F] = carry_s;
}

// src/agmv_encode.c:549-551,567-585,622-624 and :707-717 (empty audio chunk)
__global__ void __launch_bounds__(256) lz_write_chunks_k(const uint32_t* __restrict__ fs, const uint32_t* __restrict__ csize,
                                                         const uint32_t* __restrict__ chunk_off, const uint32_t* __restrict__ wbase,
                                                         const uint32_t* __restrict__ out_words, uint32_t first_frame_count,
                                                         uint8_t* __restrict__ image) {
    const uint32_t f = blockIdx.y;
    const uint32_t cs = csize[f], len = 32u + cs, usize = fs[f + 1] - fs[f];
    const uint8_t* pay = reinterpret_cast<const uint8_t*>(out_words + wbase[f]);
    uint8_t* dst = image + chunk_off[f];
    const uint32_t num = first_frame_count + f + 1;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < len; b += gridDim.x * blockDim.x) {
        uint8_t v;
        if (b < 16) {
            const uint32_t word = b < 4 ? 0x43464741u /* "AGFC" */ : (b < 8 ? num : (b < 12 ? usize : cs));
            v = (uint8_t)(word >> ((b & 3) * 8));
        } else if (b < 16 + cs) v = pay[b - 16];
        else if (b < 24 + cs) v = 0xFF;
        else {
            const uint32_t k = b - 24 - cs;
            v = k < 4 ? (uint8_t)(0x43414741u /* "AGAC" */ >> (k * 8)) : 0;
        }
        dst[b] = v;
    }
}

// Host driver. bs: batch bitstream (n bytes + >=16 bytes of readable padding);
// fs: device array of F+1 frame starts (fs[0]=0, fs[F]=n). Results: wk.csize,
// wk.outbits, wk.chunk_off on the device and the chunk image written to
// `image` (capacity >= 32*F + 9n/8 + 8).
inline void lzss_encode_batch(LzWork& wk, const uint8_t* bs, const uint32_t* fs, uint32_t F, uint32_t n,
                              uint32_t first_frame_count, uint8_t* image, LaunchCtx& lc) {
    cudaStream_t st = lc.st;
    const uint32_t nthreads = 256;
    uint32_t cover = (n + 1 > F + 1 ? n + 1 : F + 1);
    KL(lc, KC_LZ_INIT, (lz_init_k<<<cdiv(cover, nthreads), nthreads, 0, st>>>(n, fs, F, wk.A[0], wk.gs[0], wk.maxlen, wk.bestlen, wk.bitcum, wk.wbase)));
    if (n > 0) {
        for (uint32_t L = 0; L < (uint32_t)LZ_LEVELS; L++) {
            // gs[0]: group starts in A[L] order. The scatter carries them into gs[1] (A[L+1] order);
            // the running max over head flags reads gs[1] and writes the new starts back to gs[0].
            radix_pass(LzDigit{bs, wk.A[L], L}, LzMove{wk.A[L], wk.gs[0], wk.A[L + 1], wk.gs[1]}, n, wk.tile_hist, wk.scan_ws, lc);
            device_scan<MaxOp, false>(LzHead{bs, wk.A[L + 1], wk.gs[1], L},
                                      LzGroupOut{wk.A[L + 1], wk.gs[0], wk.maxlen, wk.bestlen, wk.lvlidx, wk.gsat, L + 1}, n, wk.scan_ws, lc,
                                      KC_LZ_GROUP);
        }
        uint32_t ntile = cdiv(n, PARSE_TILE);
        KL(lc, KC_LZ_PARSE, (lz_parse_spec_k<<<cdiv(ntile, 8), 256, 0, st>>>(wk.bestlen, n, ntile, wk.exit_tab, wk.w_tab)));
        KL(lc, KC_LZ_PARSE, (lz_parse_chain_k<<<1, 1024, 0, st>>>(ntile, wk.exit_tab, wk.w_tab, wk.entry_tab, wk.cumbase)));
        KL(lc, KC_LZ_PARSE, (lz_parse_mark_k<<<cdiv(ntile, 8), 256, 0, st>>>(wk.bestlen, n, ntile, wk.entry_tab, wk.cumbase, wk.bitcum)));
        size_t words = (((size_t)n * 9) >> 5) + 3 * (size_t)F + 4;
        cudaMemsetAsync(wk.out_words, 0, words * 4, st);
        APtrs ap;
        for (int l = 0; l <= LZ_LEVELS; l++) ap.a[l] = wk.A[l];
        KL(lc, KC_LZ_PACK, (lz_pack_k<<<cdiv(n, nthreads), nthreads, 0, st>>>(bs, n, fs, F, wk.bestlen, wk.lvlidx, wk.gsat, wk.bitcum, ap, wk.wbase, wk.out_words)));
    }
    KL(lc, KC_LZ_CHUNK, (lz_finalize_k<<<1, 1024, 0, st>>>(fs, F, wk.bitcum, wk.outbits, wk.csize, wk.chunk_off)));
    dim3 grid(32, F);
    KL(lc, KC_LZ_CHUNK, (lz_write_chunks_k<<<grid, 256, 0, st>>>(fs, wk.csize, wk.chunk_off, wk.wbase, wk.out_words, first_frame_count, image)));
}

}  // namespace agmvb
