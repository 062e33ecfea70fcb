// K4: exact LZSS match finding by occurrence chains (see lzchain_core.h for the algorithm and its per-position logic;
// reference: AGMV_LZSS, src/agmv_encode.c:101-177).
//
// Kernels, in launch order, for a batch of F frame bitstreams lying back to back in `bs`:
//   lzc_hashlink_k  one warp per (frame, position range): walks the range in order, 32 positions per step, with a
//                   8192-entry "last position of this 3-gram hash" table in shared memory -> hash-level links + the
//                   distance of every position to the start of its byte run. The only serial kernel of the stage: a
//                   step is one shared-memory round trip, and thousands of ranges are in flight.
//   lzc_link3_k     thread per position: hash chain -> level-3 link (most recent earlier occurrence of the 3 bytes).
//   lzc_level_k     x12, thread per position: level L -> L+1 (lzc_level). A position whose match stops growing writes its
//                   final (length, offset) on the spot.
//   (orbit.cuh)     greedy parse over bestlen[].
//   lzc_pack_k      token emission; 15-byte matches find their earliest start here (lzc_chain_end), parse-visited
//                   positions only.
// Workspace: 19 bytes per bitstream byte.
#pragma once
#include "common.cuh"
#include <algorithm>
#include <vector>
#include "lzchain_core.h"

namespace agmvb {

constexpr int LZC_HB = 13;                                 // hash bits: 8192 u32 entries = 32 KB per warp, 7 warps per SM
constexpr size_t LZC_TAB_BYTES = (size_t)4 << LZC_HB;
constexpr uint32_t LZC_PREROLL = 65536;                    // positions a range replays before its first one (>= window, multiple of 32)
constexpr int LZC_THREADS = 256;

struct LzcItem { uint32_t frame, start, end; };            // positions [start, end) of the frame (frame-relative; start a multiple of 32)

// ranges of about n / (2 x resident warps) positions, never shorter than the pre-roll (which would then dominate)
inline void lzc_build_items(const uint32_t* h_fs, uint32_t F, std::vector<LzcItem>& items) {
    items.clear();
    const uint64_t n = h_fs[F];
    uint64_t range = (n / 2072u + 31u) & ~(uint64_t)31u;
    if (range < LZC_PREROLL) range = LZC_PREROLL;
    for (uint32_t f = 0; f < F; f++) {
        const uint32_t len = h_fs[f + 1] - h_fs[f];
        for (uint64_t s = 0; s < len; s += range) items.push_back(LzcItem{f, (uint32_t)s, (uint32_t)std::min<uint64_t>(len, s + range)});
    }
}

__global__ void __launch_bounds__(32) lzc_hashlink_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs, const LzcItem* __restrict__ items,
                                                     uint32_t* __restrict__ lwh, uint16_t* __restrict__ rsd) {
    extern __shared__ uint32_t lzc_tab[];
    const LzcItem it = items[blockIdx.x];
    const uint32_t base = fs[it.frame], len = fs[it.frame + 1] - base;
    const uint32_t lane = threadIdx.x;
    for (uint32_t k = lane; k < (1u << LZC_HB); k += 32) lzc_tab[k] = 0;
    __syncwarp();
    const uint8_t* d = bs + base;
    uint32_t i0 = it.start > LZC_PREROLL ? it.start - LZC_PREROLL : 0u;
    uint32_t run_start = i0;
    // bytes i-1 .. i+2 of this lane's position, fetched one step ahead (indices clamped: the values of positions past the
    // frame end are never used)
    auto fetch = [&](uint32_t i, uint32_t& w) {
        const uint32_t a = i > 0 ? i - 1 : 0, lim = len ? len - 1 : 0;
        const uint32_t pb = d[min(a, lim)], b0 = d[min(i, lim)], b1 = d[min(i + 1, lim)], b2 = d[min(i + 2, lim)];
        w = pb | b0 << 8 | b1 << 16 | b2 << 24;
    };
    uint32_t wnext = 0;
    fetch(i0 + lane, wnext);
    for (; i0 < it.end; i0 += 32) {
        const uint32_t i = i0 + lane, w = wnext;
        if (i0 + 32 < it.end) fetch(i + 32, wnext);
        const bool valid = i + 3u <= len && i < it.end;
        const uint32_t h = lzc_hash(w >> 8, LZC_HB);
        const uint32_t old = valid ? lzc_tab[h] : 0u;
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? h : (0x80000000u | lane));
        const unsigned lower = peers & lanemask_lt();
        uint32_t dist = 0;
        if (valid) {
            const bool has = lower || old;
            const uint32_t q = lower ? i0 + (31u - (uint32_t)__clz(lower)) : old - 1u;
            if (has && i - q <= LZC_WINDOW) dist = i - q;
            if ((peers >> lane) == 1u) lzc_tab[h] = i + 1u;   // the group's last lane: most recent position of this hash
        }
        const bool brk = i < len && (i == 0 || ((w >> 8) & 0xFFu) != (w & 0xFFu));
        const unsigned bal = __ballot_sync(0xffffffffu, brk);
        const unsigned upto = bal & (0xffffffffu >> (31u - lane));
        const uint32_t rs = upto ? i0 + (31u - (uint32_t)__clz(upto)) : run_start;
        if (bal) run_start = i0 + (31u - (uint32_t)__clz(bal));
        if (i >= it.start && i < it.end) {
            lwh[base + i] = dist | (w >> 8 & 0xFFFFu) << 16;
            rsd[base + i] = (uint16_t)min(i - rs, 65535u);
        }
        __syncwarp();
    }
}

// largest f with fs[f] <= i and fs[f+1] > i (empty frames are skipped); i < fs[F]
__device__ __forceinline__ uint32_t lzc_frame_of(const uint32_t* __restrict__ fs, uint32_t F, uint32_t i) {
    uint32_t lo = 0, hi = F;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (fs[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}
// bytes left in p's frame; the block's first position is searched once, the others walk forward from it
__device__ __forceinline__ uint32_t lzc_rem(const uint32_t* __restrict__ fs, uint32_t F, uint32_t p, uint32_t n) {
    __shared__ uint32_t s_f;
    if (threadIdx.x == 0) s_f = lzc_frame_of(fs, F, min(blockIdx.x * blockDim.x, n - 1));
    __syncthreads();
    if (p >= n) return 0;
    uint32_t f = s_f;
    while (fs[f + 1] <= p) f++;
    return fs[f + 1] - p;
}

__global__ void __launch_bounds__(LZC_THREADS) lzc_link3_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs, uint32_t F, uint32_t n,
                                                           const uint32_t* __restrict__ lwh, const uint16_t* __restrict__ rsd,
                                                           uint32_t* __restrict__ lw3, uint8_t* __restrict__ bestlen) {
    const uint32_t p = blockIdx.x * LZC_THREADS + threadIdx.x;
    const uint32_t rem = lzc_rem(fs, F, p, n);
    if (p >= n) return;
    const uint32_t d3 = lzc_link3(bs, lwh, rsd, p, rem);
    lw3[p] = d3 | (uint32_t)bs[p + 3] << 16;
    if (!d3) bestlen[p] = 0;
}

__global__ void __launch_bounds__(LZC_THREADS) lzc_level_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs, uint32_t F, uint32_t n, uint32_t L,
                                                           const uint32_t* __restrict__ lw, const uint16_t* __restrict__ rsd, uint32_t* __restrict__ lw_next,
                                                           uint32_t* __restrict__ match_rec, uint8_t* __restrict__ bestlen) {
    const uint32_t p = blockIdx.x * LZC_THREADS + threadIdx.x;
    const uint32_t rem = lzc_rem(fs, F, p, n);
    if (p >= n) return;
    uint32_t rec = 0;
    const uint32_t nd = lzc_level(bs, lw, rsd, p, rem, L, &rec);
    lw_next[p] = nd | (uint32_t)bs[p + L + 1] << 16;
    if (rec) { match_rec[p] = rec; bestlen[p] = (uint8_t)L; }
    else if (nd && L + 1 == (uint32_t)LZ_MAXLEN) bestlen[p] = (uint8_t)LZ_MAXLEN;
}

__global__ void lzc_wbase_k(const uint32_t* __restrict__ fs, uint32_t F, uint32_t* __restrict__ wbase) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= F) wbase[i] = (uint32_t)(((uint64_t)fs[i] * 9u) >> 5) + 3u * i;
}

// token emission (bit writer: src/agmv_utils.c:86-112; token layout src/agmv_encode.c:146-165)
__global__ void __launch_bounds__(256) lzc_pack_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs, const uint8_t* __restrict__ bestlen,
                                                  const uint32_t* __restrict__ match_rec, const uint32_t* __restrict__ bitcum,
                                                  const uint32_t* __restrict__ lw15, const uint16_t* __restrict__ rsd,
                                                  const uint32_t* __restrict__ wbase, uint32_t* __restrict__ out_words) {
    const uint32_t f = blockIdx.y;  // one grid row per frame: no search for the frame of a position
    const uint32_t i = fs[f] + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fs[f + 1]) return;
    const uint32_t rel = bitcum[i];
    if (rel == EMPTY32) return;
    const uint32_t l = bestlen[i];
    uint32_t v, nb;
    if (l >= (uint32_t)LZ_MINLEN) {
        const uint32_t off = l == (uint32_t)LZ_MAXLEN ? lzc_chain_end(lw15, rsd, i) : match_rec[i] & 0xFFFFu;
        v = (off << 1) | (l << 17);
        nb = 21;
    } else {
        v = 1u | ((uint32_t)bs[i] << 1);
        nb = 9;
    }
    const uint32_t w = wbase[f] + (rel >> 5), sh = rel & 31;
    atomicOr(&out_words[w], v << sh);
    if (sh + nb > 32) atomicOr(&out_words[w + 1], v >> (32 - sh));
}

}  // namespace agmvb
