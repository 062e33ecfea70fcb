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

// ---- the walking kernels ---------------------------------------------------------------------------------------
// A walk is a chain of dependent gathers whose length varies from position to position (most end at the first hop, a
// few per cent must cross their whole window), so "one thread walks one position to the end" leaves most lanes of a warp
// waiting for the slowest one. Both kernels therefore work in two phases on a warp's 512 consecutive positions:
//   1. a coalesced sweep that takes every position's FIRST hop (four independent rounds in flight) and finishes the
//      positions it settles; the others are queued (shared memory, 2 B each);
//   2. the queue is drained with lane refill: a lane whose walk ends takes the next queued position at once, so every
//      hop instruction has (almost) all lanes doing useful work.
constexpr int LZC_ROUNDS = 16;
constexpr int LZC_WCHUNK = 32 * LZC_ROUNDS;           // positions per warp
constexpr int LZC_WARPS = LZC_THREADS / 32;
constexpr int LZC_BCHUNK = LZC_WCHUNK * LZC_WARPS;    // positions per block
constexpr int LZC_MLP = 4;                            // rounds of phase 1 in flight together

__global__ void __launch_bounds__(LZC_THREADS) lzc_link3_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs, uint32_t F, uint32_t n,
                                                           const uint32_t* __restrict__ lwh, const uint16_t* __restrict__ rsd,
                                                           uint32_t* __restrict__ lw3, uint8_t* __restrict__ bestlen) {
    __shared__ uint16_t q[LZC_WARPS][LZC_WCHUNK];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t cbase = blockIdx.x * LZC_BCHUNK + warp * LZC_WCHUNK;
    if (cbase >= n) return;
    uint32_t f0 = 0;
    if (lane == 0) f0 = lzc_frame_of(fs, F, cbase);
    f0 = __shfl_sync(0xffffffffu, f0, 0);
    auto cap_of = [&](uint32_t p) {   // min(15, bytes left in p's frame)
        uint32_t f = f0;
        while (fs[f + 1] <= p) f++;
        return min(fs[f + 1] - p, (uint32_t)LZ_MAXLEN);
    };
    uint32_t qn = 0;
    for (int r0 = 0; r0 < LZC_ROUNDS; r0 += LZC_MLP) {
        uint32_t w[LZC_MLP], b23[LZC_MLP], cap[LZC_MLP], wk[LZC_MLP], k2[LZC_MLP];
#pragma unroll
        for (int j = 0; j < LZC_MLP; j++) {
            const uint32_t p = cbase + (r0 + j) * 32 + lane;
            w[j] = 0; b23[j] = 0; cap[j] = 0;
            if (p < n) { w[j] = lwh[p]; b23[j] = (uint32_t)bs[p + 2] | (uint32_t)bs[p + 3] << 8; cap[j] = cap_of(p); }
        }
#pragma unroll
        for (int j = 0; j < LZC_MLP; j++) {
            const uint32_t p = cbase + (r0 + j) * 32 + lane, dist = w[j] & 0xFFFFu;
            wk[j] = 0; k2[j] = 0;
            if (cap[j] >= 3u && dist) { wk[j] = lwh[p - dist]; k2[j] = bs[p - dist + 2]; }
        }
#pragma unroll
        for (int j = 0; j < LZC_MLP; j++) {
            const uint32_t p = cbase + (r0 + j) * 32 + lane, dist = w[j] & 0xFFFFu;
            bool pend = false;
            if (p < n) {
                uint32_t nd = 0;
                if (cap[j] >= 3u && dist) {
                    if ((wk[j] >> 16) == (w[j] >> 16) && k2[j] == (b23[j] & 0xFFu)) nd = dist;
                    else pend = true;
                }
                if (!pend) {
                    lw3[p] = lzc_word(nd, b23[j] >> 8, b23[j] & 0xFFu, cap[j]);
                    if (!nd) bestlen[p] = 0;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pend);
            if (pend) q[warp][qn + __popc(bal & lanemask_lt())] = (uint16_t)((r0 + j) * 32 + lane);
            qn += __popc(bal);
        }
    }
    __syncwarp();
    uint32_t qi = 0, b23c = 0;
    bool busy = false;
    LzcLink3Walk wlk;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (qi < qn && idle) {
            const uint32_t my = qi + __popc(idle & lanemask_lt());
            if (!busy && my < qn) {
                const uint32_t p = cbase + q[warp][my];
                const uint32_t cp = cap_of(p);
                b23c = (uint32_t)bs[p + 2] | (uint32_t)bs[p + 3] << 8 | cp << 16;
                wlk.start(p, lwh[p], b23c & 0xFFu, cp);
                busy = true;
            }
            qi += __popc(idle);
        }
        if (!__any_sync(0xffffffffu, busy)) break;
        if (busy) {
            const int r = wlk.hop(bs, lwh, rsd);
            if (r != LZC_GO) {
                const uint32_t nd = r == LZC_FOUND ? wlk.acc : 0u;
                lw3[wlk.p] = lzc_word(nd, (b23c >> 8) & 0xFFu, b23c & 0xFFu, b23c >> 16);
                if (!nd) bestlen[wlk.p] = 0;
                busy = false;
            }
        }
    }
}

__global__ void __launch_bounds__(LZC_THREADS) lzc_level_k(const uint8_t* __restrict__ bs, uint32_t n, uint32_t L, const uint32_t* __restrict__ lw,
                                                           const uint16_t* __restrict__ rsd, uint32_t* __restrict__ lw_next,
                                                           uint32_t* __restrict__ match_rec, uint8_t* __restrict__ bestlen) {
    __shared__ uint16_t q[LZC_WARPS][LZC_WCHUNK];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t cbase = blockIdx.x * LZC_BCHUNK + warp * LZC_WCHUNK;
    if (cbase >= n) return;
    const bool top = L + 1u == (uint32_t)LZ_MAXLEN;
    uint32_t qn = 0;
    for (int r0 = 0; r0 < LZC_ROUNDS; r0 += LZC_MLP) {
        uint32_t w[LZC_MLP], nb[LZC_MLP], wk[LZC_MLP];
#pragma unroll
        for (int j = 0; j < LZC_MLP; j++) {
            const uint32_t p = cbase + (r0 + j) * 32 + lane;
            w[j] = 0; nb[j] = 0;
            if (p < n) { w[j] = lw[p]; nb[j] = bs[p + L + 1]; }
        }
#pragma unroll
        for (int j = 0; j < LZC_MLP; j++) {
            const uint32_t p = cbase + (r0 + j) * 32 + lane, dist = w[j] & 0xFFFFu;
            wk[j] = 0;
            if (dist) wk[j] = lw[p - dist];
        }
#pragma unroll
        for (int j = 0; j < LZC_MLP; j++) {
            const uint32_t p = cbase + (r0 + j) * 32 + lane, dist = w[j] & 0xFFFFu;
            const uint32_t c = (w[j] >> 16) & 0xFFu, cap = (w[j] >> 24) & 0xFu;
            bool pend = false;
            if (p < n) {
                uint32_t nd = 0;
                if (dist) {
                    if (L + 1u <= cap && ((wk[j] >> 16) & 0xFFu) == c) nd = dist;
                    else pend = true;
                }
                if (!pend) {
                    lw_next[p] = lzc_word(nd, nb[j], c, cap);
                    if (nd && top) bestlen[p] = (uint8_t)LZ_MAXLEN;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pend);
            if (pend) q[warp][qn + __popc(bal & lanemask_lt())] = (uint16_t)((r0 + j) * 32 + lane);
            qn += __popc(bal);
        }
    }
    __syncwarp();
    uint32_t qi = 0, nbc = 0;
    bool busy = false;
    LzcLevelWalk wlk;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (qi < qn && idle) {
            const uint32_t my = qi + __popc(idle & lanemask_lt());
            if (!busy && my < qn) {
                const uint32_t p = cbase + q[warp][my];
                const uint32_t w = lw[p];
                nbc = (uint32_t)bs[p + L + 1] | ((w >> 24) & 0xFu) << 8;
                wlk.start(p, w, L);
                busy = true;
            }
            qi += __popc(idle);
        }
        if (!__any_sync(0xffffffffu, busy)) break;
        if (busy) {
            const int r = wlk.hop(lw, rsd);
            if (r != LZC_GO) {
                const uint32_t nd = r == LZC_FOUND ? wlk.acc : 0u;
                lw_next[wlk.p] = lzc_word(nd, nbc & 0xFFu, wlk.c, nbc >> 8);
                if (r == LZC_END) { match_rec[wlk.p] = L << 28 | LZC_RESOLVED | wlk.last; bestlen[wlk.p] = (uint8_t)L; }
                else if (top) bestlen[wlk.p] = (uint8_t)LZ_MAXLEN;
                busy = false;
            }
        }
    }
}

__global__ void lzc_wbase_k(const uint32_t* __restrict__ fs, uint32_t F, uint32_t* __restrict__ wbase) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= F) wbase[i] = (uint32_t)(((uint64_t)fs[i] * 9u) >> 5) + 3u * i;
}

// token emission (bit writer: src/agmv_utils.c:86-112; token layout src/agmv_encode.c:146-165). Same two phases as the
// walking kernels: literals and matches shorter than 15 bytes are emitted in the sweep, the 15-byte matches of the parse
// (the only positions whose earliest start is still unknown) walk their level-15 chain to its end with lane refill.
__device__ __forceinline__ void lzc_emit(uint32_t* __restrict__ out_words, uint32_t wb, uint32_t rel, uint32_t v, uint32_t nb) {
    const uint32_t w = wb + (rel >> 5), sh = rel & 31;
    atomicOr(&out_words[w], v << sh);
    if (sh + nb > 32) atomicOr(&out_words[w + 1], v >> (32 - sh));
}
__global__ void __launch_bounds__(LZC_THREADS) lzc_pack_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs, const uint8_t* __restrict__ bestlen,
                                                          const uint32_t* __restrict__ match_rec, const uint32_t* __restrict__ bitcum,
                                                          const uint32_t* __restrict__ lw15, const uint16_t* __restrict__ rsd,
                                                          const uint32_t* __restrict__ wbase, uint32_t* __restrict__ out_words) {
    __shared__ uint16_t q[LZC_WARPS][LZC_WCHUNK];
    const uint32_t f = blockIdx.y;  // one grid row per frame: no search for the frame of a position
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t cbase = fs[f] + blockIdx.x * LZC_BCHUNK + warp * LZC_WCHUNK, end = fs[f + 1];
    if (cbase >= end) return;
    const uint32_t wb = wbase[f];
    uint32_t qn = 0;
#pragma unroll 4
    for (int r = 0; r < LZC_ROUNDS; r++) {
        const uint32_t i = cbase + r * 32 + lane;
        bool pend = false;
        if (i < end) {
            const uint32_t rel = bitcum[i];
            if (rel != EMPTY32) {
                const uint32_t l = bestlen[i];
                if (l == (uint32_t)LZ_MAXLEN) pend = true;
                else if (l >= (uint32_t)LZ_MINLEN) lzc_emit(out_words, wb, rel, ((match_rec[i] & 0xFFFFu) << 1) | (l << 17), 21);
                else lzc_emit(out_words, wb, rel, 1u | ((uint32_t)bs[i] << 1), 9);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pend);
        if (pend) q[warp][qn + __popc(bal & lanemask_lt())] = (uint16_t)(r * 32 + lane);
        qn += __popc(bal);
    }
    __syncwarp();
    uint32_t qi = 0;
    bool busy = false;
    LzcEndWalk wlk;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (qi < qn && idle) {
            const uint32_t my = qi + __popc(idle & lanemask_lt());
            if (!busy && my < qn) {
                const uint32_t p = cbase + q[warp][my];
                wlk.start(p, lw15[p]);   // a 15-byte match: the link is never 0
                busy = true;
            }
            qi += __popc(idle);
        }
        if (!__any_sync(0xffffffffu, busy)) break;
        if (busy && wlk.hop(lw15, rsd) != LZC_GO) {
            lzc_emit(out_words, wb, bitcum[wlk.p], (wlk.last << 1) | ((uint32_t)LZ_MAXLEN << 17), 21);
            busy = false;
        }
    }
}

}  // namespace agmvb
