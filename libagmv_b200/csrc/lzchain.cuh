// K4: exact LZSS match finding by occurrence chains (see lzchain_core.h for the algorithm and its per-position logic;
// reference: AGMV_LZSS, src/agmv_encode.c:101-177).
//
// Kernels, in launch order, for a batch of F frame bitstreams lying back to back in `bs`:
//   lzc_hashlink_k  one warp per (frame, position range): walks the range in order, 32 positions per step, with a
//                   8192-entry "last position of this 3-gram hash" table in shared memory -> hash-level links + the
//                   distance of every position to the start of its byte run. The only serial kernel of the stage: a
//                   step is one shared-memory round trip, and thousands of ranges are in flight.
//   lzt_pass_k<0>   hash chain -> level-3 link (most recent earlier occurrence of the 3 bytes).
//   lzt_pass_k<1>   x12: level L -> L+1. A position whose match stops growing writes its final (length, offset) on the spot.
//                   Both are the tiled walker below: every gather of a chain walk is served from shared memory.
//   (orbit.cuh)     greedy parse over bestlen[].
//   lzc_pack_k      token emission; 15-byte matches find their earliest start here (end of their level-15 chain),
//                   parse-visited positions only.
// Workspace: 15 bytes per bitstream byte + 4.9 MB per resident block of the tiled kernel.
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

// The bytes of 16 steps (512 positions + 3) are staged in shared memory with coalesced word loads issued one block ahead,
// so the serial loop never waits for global memory: a step is stage read -> hash -> table read -> match -> table write.
constexpr int LZC_STAGE_STEPS = 16;
constexpr int LZC_STAGE_WORDS = 32 * 5;   // 640 bytes >= 32 * 16 + 3 + 3 bytes of misalignment

__global__ void __launch_bounds__(32) lzc_hashlink_k(const uint8_t* __restrict__ bs, uint32_t n, const uint32_t* __restrict__ fs,
                                                     const LzcItem* __restrict__ items, uint16_t* __restrict__ dh, uint16_t* __restrict__ rsd) {
    extern __shared__ uint32_t lzc_tab[];
    __shared__ uint32_t stage[2][LZC_STAGE_WORDS + 1];
    const LzcItem it = items[blockIdx.x];
    const uint32_t base = fs[it.frame], len = fs[it.frame + 1] - base;
    const uint32_t lane = threadIdx.x;
    for (uint32_t k = lane; k < (1u << LZC_HB); k += 32) lzc_tab[k] = 0;
    uint32_t i0 = it.start > LZC_PREROLL ? it.start - LZC_PREROLL : 0u;
    uint32_t run_start = i0;
    // readable words of the batch buffer: [word_lo, word_hi] (bs has 16 bytes of padding after its n bytes)
    const uintptr_t word_lo = reinterpret_cast<uintptr_t>(bs) & ~(uintptr_t)3;
    const uintptr_t word_hi = (reinterpret_cast<uintptr_t>(bs) + n + 12u) & ~(uintptr_t)3;
    // block of 16 steps starting at position j0: words from the aligned address at or below byte j0 - 1
    uint32_t regs[5];
    auto block_addr = [&](uint32_t j0) {
        const uintptr_t want = reinterpret_cast<uintptr_t>(bs) + base + j0 - 1u;   // byte before the block's first position
        const uintptr_t a = want & ~(uintptr_t)3;
        return a < word_lo ? word_lo : a;
    };
    auto fetch = [&](uint32_t j0) {
        const uintptr_t a = block_addr(j0);
#pragma unroll
        for (int k = 0; k < 5; k++) {
            uintptr_t x = a + 4u * (uint32_t)(k * 32 + lane);
            if (x > word_hi) x = word_hi;
            regs[k] = *reinterpret_cast<const uint32_t*>(x);
        }
    };
    fetch(i0);
    int buf = 0;
    for (uint32_t j0 = i0; j0 < it.end; j0 += 32 * LZC_STAGE_STEPS, buf ^= 1) {
#pragma unroll
        for (int k = 0; k < 5; k++) stage[buf][k * 32 + lane] = regs[k];
        if (lane == 0) stage[buf][LZC_STAGE_WORDS] = 0;
        __syncwarp();
        if (j0 + 32 * LZC_STAGE_STEPS < it.end) fetch(j0 + 32 * LZC_STAGE_STEPS);
        // byte offset of position j0 - 1 inside the staged words (-1 only for the very first byte of the buffer)
        const int off0 = (int)(reinterpret_cast<uintptr_t>(bs) + base + j0 - 1u - block_addr(j0));
        const uint32_t jend = min(j0 + 32 * LZC_STAGE_STEPS, it.end);
        for (uint32_t s0 = j0; s0 < jend; s0 += 32) {
            const uint32_t i = s0 + lane;
            const int o = off0 + (int)(i - j0);
            // bytes i-1 .. i+2 (o == -1: the buffer's very first position, which has no byte before it)
            const uint32_t w = o < 0 ? stage[buf][0] << 8 : __funnelshift_r(stage[buf][o >> 2], stage[buf][(o >> 2) + 1], (o & 3) * 8);
            const bool valid = i + 3u <= len && i < it.end;
            const uint32_t h = lzc_hash(w >> 8, LZC_HB);
            const uint32_t old = valid ? lzc_tab[h] : 0u;
            const unsigned peers = __match_any_sync(0xffffffffu, valid ? h : (0x80000000u | lane));
            const unsigned lower = peers & lanemask_lt();
            uint32_t dist = 0;
            if (valid) {
                const bool has = lower || old;
                const uint32_t q = lower ? s0 + (31u - (uint32_t)__clz(lower)) : old - 1u;
                if (has && i - q <= LZC_WINDOW) dist = i - q;
                if ((peers >> lane) == 1u) lzc_tab[h] = i + 1u;   // the group's last lane: most recent position of this hash
            }
            const bool brk = i < len && (i == 0 || ((w >> 8) & 0xFFu) != (w & 0xFFu));
            const unsigned bal = __ballot_sync(0xffffffffu, brk);
            const unsigned upto = bal & (0xffffffffu >> (31u - lane));
            const uint32_t rs = upto ? s0 + (31u - (uint32_t)__clz(upto)) : run_start;
            if (bal) run_start = s0 + (31u - (uint32_t)__clz(bal));
            if (i >= it.start && i < it.end) {
                dh[base + i] = (uint16_t)dist;
                rsd[base + i] = (uint16_t)min(i - rs, 65535u);
            }
            __syncwarp();
        }
    }
}

// ---- the tiled walker ------------------------------------------------------------------------------------------
// A chain walk is a string of dependent 2-byte gathers into the 64 KB behind a position. From global memory such gathers
// cost an L1 wavefront per lane (measured: the level kernels ran at exactly that limit, ~2 cycles per lane-gather); from
// shared memory they are nearly free. So a block takes one frame at a time and visits its 16384-position tiles from the
// LAST to the FIRST, holding only the current tile (links + bytes, 48 KB) in shared memory:
//   * sweep: every position of the tile starts its walk. A first hop that stays inside the tile is checked on the spot
//     (97 % of the first hops do, and most end the walk); otherwise the walker (12 bytes) is parked in the bin of the tile
//     its next hop lands in;
//   * drain: the walkers parked in this tile's bin - the tile's own and those that came down from the four tiles above -
//     hop inside the tile until they end or leave it downwards (parked again, in a lower bin). Lanes refill from the bin as
//     their walks end, so walks of very different lengths keep the warp full.
// Because tiles go downwards, a bin is complete when its tile becomes resident, and every tile is loaded exactly once per
// level. The bins live in global memory (L2): 5 bins per block, each sized for the worst case (every position of the five
// tiles that can reach it). Four blocks are resident per SM, so one block's barrier waits hide behind the others' work.
constexpr int LZT_LOG = 14, LZT_TILE = 1 << LZT_LOG;
constexpr int LZT_BINS = 5;                                   // resident tile + the 4 below it (4 * 16384 >= 65535)
constexpr int LZT_THREADS = 256;
constexpr uint32_t LZT_BINCAP = (uint32_t)LZT_BINS * LZT_TILE;
constexpr size_t LZT_ARENA_WORDS = (size_t)LZT_BINS * LZT_BINCAP * 3;   // per block: 3 words per parked walker
constexpr int LZT_RAW = LZT_TILE + 32;                        // bytes of a tile + what a hop reads past its last position
struct LztSmem {
    uint4 sd4[LZT_TILE / 8 + 1];     // links of the tile (u16), shifted so that 16-byte chunks of the global array stay aligned
    uint4 sraw4[LZT_RAW / 16 + 2];   // bytes of the tile, shifted likewise
    uint32_t cnt[LZT_BINS];          // walkers parked per bin slot
    uint32_t head;                   // next walker of the bin being drained
    uint32_t frame;
};

// park walkers: lanes with `push` append w to bin slot `bin` (aggregated per bin: one shared-memory atomic per bin and warp)
__device__ __forceinline__ void lzt_push(LztSmem& S, uint32_t* __restrict__ arena, bool push, uint32_t bin, const LzcWalker& w) {
    if (!__ballot_sync(0xffffffffu, push)) return;
    uint32_t slot = 0;
#pragma unroll
    for (uint32_t b = 0; b < (uint32_t)LZT_BINS; b++) {
        const unsigned m = __ballot_sync(0xffffffffu, push && bin == b);
        if (m) {
            const int leader = __ffs(m) - 1;
            uint32_t b0 = 0;
            if ((int)lane_id() == leader) b0 = atomicAdd(&S.cnt[b], (uint32_t)__popc(m));
            b0 = __shfl_sync(0xffffffffu, b0, leader);
            if (push && bin == b) slot = b0 + __popc(m & lanemask_lt());
        }
    }
    if (push && slot < LZT_BINCAP) {
        uint32_t* e = arena + ((size_t)bin * LZT_BINCAP + slot) * 3;
        e[0] = w.p;
        e[1] = w.acc | w.dist << 16;
        e[2] = w.key;
    }
}

// copy `bytes` bytes from global `src` into shared memory in aligned 16-byte chunks; returns the byte offset of src[0] inside
// dst (src & 15). May read up to 15 bytes before src and 15 after src + bytes (inside the same allocations).
__device__ __forceinline__ uint32_t lzt_load(uint4* dst, const void* src, uint32_t bytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uint32_t shift = (uint32_t)(a & 15u);
    const uint4* g = reinterpret_cast<const uint4*>(a - shift);
    const uint32_t chunks = (shift + bytes + 15u) >> 4;
    for (uint32_t c = threadIdx.x; c < chunks; c += blockDim.x) dst[c] = g[c];
    return shift;
}

// MODE 0: hash links -> level-3 links (L unused). MODE 1: level L -> L+1.
template <int MODE>
__global__ void __launch_bounds__(LZT_THREADS) lzt_pass_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs, uint32_t F, uint32_t L,
                                                          const uint16_t* __restrict__ din, const uint16_t* __restrict__ rsd, uint16_t* __restrict__ dout,
                                                          uint32_t* __restrict__ match_rec, uint8_t* __restrict__ bestlen,
                                                          uint32_t* __restrict__ arena, uint32_t* __restrict__ counter) {
    extern __shared__ __align__(16) uint8_t lzt_smem_raw[];
    LztSmem& S = *reinterpret_cast<LztSmem*>(lzt_smem_raw);
    uint32_t* const my_arena = arena + (size_t)blockIdx.x * LZT_ARENA_WORDS;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const bool top = MODE == 1 && L + 1u == (uint32_t)LZ_MAXLEN;
    for (;;) {
        __syncthreads();
        if (tid == 0) S.frame = atomicAdd(counter, 1u);
        if (tid < (uint32_t)LZT_BINS) S.cnt[tid] = 0;
        __syncthreads();
        const uint32_t f = S.frame;
        if (f >= F) break;
        const uint32_t base = fs[f], len = fs[f + 1] - base;
        const uint16_t* const g_in = din + base;
        const uint16_t* const g_rsd = rsd + base;
        const uint8_t* const raw = bs + base;
        for (uint32_t j = (len + LZT_TILE - 1) >> LZT_LOG; j-- > 0;) {
            const uint32_t t0 = j << LZT_LOG, cntp = min(len - t0, (uint32_t)LZT_TILE), slot = j % (uint32_t)LZT_BINS;
            const uint32_t sh_d = lzt_load(S.sd4, g_in + t0, cntp * 2u);
            const uint32_t sh_r = lzt_load(S.sraw4, raw + t0, cntp + 32u);   // bs is readable 64 bytes past the batch
            if (tid == 0) S.head = 0;
            __syncthreads();
            const uint16_t* const sd = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(S.sd4) + sh_d);
            const uint8_t* const sraw = reinterpret_cast<const uint8_t*>(S.sraw4) + sh_r;
            // ---- sweep ----
            for (uint32_t i0 = 0; i0 < cntp; i0 += LZT_THREADS) {
                const uint32_t i = i0 + tid;
                bool push = false;
                uint32_t bin = 0;
                LzcWalker w{0u, 0u, 0u, 0u};
                if (i < cntp) {
                    const uint32_t p = t0 + i, dist = sd[i], cap = min(len - p, (uint32_t)LZ_MAXLEN);
                    if (!dist || (MODE == 0 && cap < (uint32_t)LZ_MINLEN)) {
                        dout[base + p] = 0;
                        if (MODE == 0) bestlen[base + p] = 0;
                    } else {
                        w.p = p;
                        w.dist = dist;
                        if (MODE == 0) w.key = (uint32_t)sraw[i] | (uint32_t)sraw[i + 1] << 8 | (uint32_t)sraw[i + 2] << 16;
                        else w.key = lzc_level_key(sraw[i + L], sraw[i + L - 1], L + 1u <= cap);
                        const uint32_t k = p - dist;
                        if (k >= t0) {   // first hop inside the tile: settle it here if it ends the walk
                            const uint32_t kk = k - t0;
                            bool hit;
                            if (MODE == 0) hit = ((uint32_t)sraw[kk] | (uint32_t)sraw[kk + 1] << 8 | (uint32_t)sraw[kk + 2] << 16) == w.key;
                            else hit = ((w.key >> 8) & 1u) && sraw[kk + L] == (w.key & 0xFFu);
                            if (hit) {
                                dout[base + p] = (uint16_t)dist;
                                if (top) bestlen[base + p] = (uint8_t)LZ_MAXLEN;
                            } else { push = true; bin = slot; }
                        } else { push = true; bin = (k >> LZT_LOG) % (uint32_t)LZT_BINS; }
                    }
                }
                lzt_push(S, my_arena, push, bin, w);
            }
            __syncthreads();
            // ---- drain ----
            const uint32_t total = min(S.cnt[slot], LZT_BINCAP);
            const uint32_t* const qa = my_arena + (size_t)slot * LZT_BINCAP * 3;
            bool busy = false, dry = false;
            LzcWalker w{0u, 0u, 0u, 0u};
            for (;;) {
                const unsigned idle = __ballot_sync(0xffffffffu, !busy);
                if (idle && !dry) {
                    const int leader = __ffs(idle) - 1;
                    uint32_t b0 = 0;
                    if ((int)lane == leader) b0 = atomicAdd(&S.head, (uint32_t)__popc(idle));
                    b0 = __shfl_sync(0xffffffffu, b0, leader);
                    const uint32_t e = b0 + __popc(idle & lanemask_lt());
                    if (!busy && e < total) {
                        const uint32_t* q = qa + (size_t)e * 3;
                        const uint32_t x = q[1];
                        w.p = q[0]; w.acc = x & 0xFFFFu; w.dist = x >> 16; w.key = q[2];
                        busy = true;
                    }
                    dry = b0 + (uint32_t)__popc(idle) >= total;
                }
                if (!__any_sync(0xffffffffu, busy)) break;
                bool push = false;
                uint32_t bin = 0;
                if (busy) {
                    const int r = MODE == 0 ? lzc_link3_hop(w, t0, sd, sraw, g_in, g_rsd) : lzc_level_hop(w, L, t0, sd, sraw, g_in, g_rsd);
                    if (r == LZC_FOUND) {
                        dout[base + w.p] = (uint16_t)w.acc;
                        if (top) bestlen[base + w.p] = (uint8_t)LZ_MAXLEN;
                        busy = false;
                    } else if (r == LZC_END) {
                        dout[base + w.p] = 0;
                        if (MODE == 0) bestlen[base + w.p] = 0;
                        else { match_rec[base + w.p] = L << 28 | LZC_RESOLVED | w.acc; bestlen[base + w.p] = (uint8_t)L; }
                        busy = false;
                    } else if (r == LZC_LEAVE) {
                        push = true;
                        bin = ((w.p - w.acc - w.dist) >> LZT_LOG) % (uint32_t)LZT_BINS;
                        busy = false;
                    }
                }
                lzt_push(S, my_arena, push, bin, w);
            }
            __syncthreads();
            if (tid == 0) S.cnt[slot] = 0;
        }
    }
}

__global__ void lzc_wbase_k(const uint32_t* __restrict__ fs, uint32_t F, uint32_t* __restrict__ wbase) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= F) wbase[i] = (uint32_t)(((uint64_t)fs[i] * 9u) >> 5) + 3u * i;
}

// token emission (bit writer: src/agmv_utils.c:86-112; token layout src/agmv_encode.c:146-165). Literals and matches
// shorter than 15 bytes are emitted in a coalesced sweep; the 15-byte matches of the parse (the only positions whose earliest
// start is still unknown) walk their level-15 chain to its end, queued per warp and drained with lane refill.
constexpr int LZC_ROUNDS = 16;
constexpr int LZC_WCHUNK = 32 * LZC_ROUNDS;           // positions per warp
constexpr int LZC_WARPS = LZC_THREADS / 32;
constexpr int LZC_BCHUNK = LZC_WCHUNK * LZC_WARPS;    // positions per block
__device__ __forceinline__ void lzc_emit(uint32_t* __restrict__ out_words, uint32_t wb, uint32_t rel, uint32_t v, uint32_t nb) {
    const uint32_t w = wb + (rel >> 5), sh = rel & 31;
    atomicOr(&out_words[w], v << sh);
    if (sh + nb > 32) atomicOr(&out_words[w + 1], v >> (32 - sh));
}
__global__ void __launch_bounds__(LZC_THREADS) lzc_pack_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs, const uint8_t* __restrict__ bestlen,
                                                          const uint32_t* __restrict__ match_rec, const uint32_t* __restrict__ bitcum,
                                                          const uint16_t* __restrict__ d15, const uint16_t* __restrict__ rsd,
                                                          const uint32_t* __restrict__ wbase, uint32_t* __restrict__ out_words) {
    __shared__ uint16_t q[LZC_WARPS][LZC_WCHUNK];
    const uint32_t f = blockIdx.y;  // one grid row per frame: no search for the frame of a position
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t fbase = fs[f], cbase = blockIdx.x * LZC_BCHUNK + warp * LZC_WCHUNK, flen = fs[f + 1] - fbase;   // frame-relative chunk
    if (cbase >= flen) return;
    const uint32_t wb = wbase[f];
    const uint32_t* const bc = bitcum + fbase;
    uint32_t qn = 0;
#pragma unroll 4
    for (int r = 0; r < LZC_ROUNDS; r++) {
        const uint32_t i = cbase + r * 32 + lane;
        bool pend = false;
        if (i < flen) {
            const uint32_t rel = bc[i];
            if (rel != EMPTY32) {
                const uint32_t l = bestlen[fbase + i];
                if (l == (uint32_t)LZ_MAXLEN) pend = true;
                else if (l >= (uint32_t)LZ_MINLEN) lzc_emit(out_words, wb, rel, ((match_rec[fbase + i] & 0xFFFFu) << 1) | (l << 17), 21);
                else lzc_emit(out_words, wb, rel, 1u | ((uint32_t)bs[fbase + i] << 1), 9);
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pend);
        if (pend) q[warp][qn + __popc(bal & lanemask_lt())] = (uint16_t)(r * 32 + lane);
        qn += __popc(bal);
    }
    __syncwarp();
    uint32_t qi = 0;
    bool busy = false;
    LzcWalker w{0u, 0u, 0u, 0u};
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (qi < qn && idle) {
            const uint32_t my = qi + __popc(idle & lanemask_lt());
            if (!busy && my < qn) {
                w.p = cbase + q[warp][my];
                w.acc = 0;
                w.dist = d15[fbase + w.p];   // a 15-byte match: the link is never 0
                busy = true;
            }
            qi += __popc(idle);
        }
        if (!__any_sync(0xffffffffu, busy)) break;
        if (busy && lzc_end_hop(w, d15 + fbase, rsd + fbase) != LZC_GO) {
            lzc_emit(out_words, wb, bc[w.p], (w.acc << 1) | ((uint32_t)LZ_MAXLEN << 17), 21);
            busy = false;
        }
    }
}

}  // namespace agmvb
