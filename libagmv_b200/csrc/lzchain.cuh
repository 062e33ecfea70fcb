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
//   (orbit.cuh, lzss.cuh)  greedy parse over bestlen[]; its last pass emits the tokens (lz_emit_mark_k) except the 15-byte
//                   matches, which find their earliest start by a chain walk (lz_pack15_k), parse-visited positions only.
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
inline void lzc_build_items(const uint32_t* h_fs, uint32_t F, std::vector<LzcItem>& items, int hash_bits = LZC_HB) {
    items.clear();
    const uint64_t n = h_fs[F];
    const uint64_t resident = 148ull * std::min<uint64_t>(32, (227u << 10) / (((size_t)4 << hash_bits) + 2200));   // one-warp blocks the GPU holds
    uint64_t range = (n / (2 * resident) + 31u) & ~(uint64_t)31u;   // two waves
    if (range < LZC_PREROLL) range = LZC_PREROLL;
    for (uint32_t f = 0; f < F; f++) {
        const uint32_t len = h_fs[f + 1] - h_fs[f];
        for (uint64_t s = 0; s < len; s += range) items.push_back(LzcItem{f, (uint32_t)s, (uint32_t)std::min<uint64_t>(len, s + range)});
    }
}

// The bytes of 16 steps (512 positions + 3) are staged in shared memory by the TMA engine: one elected lane issues a 1-D
// bulk copy (cp.async.bulk, completion counted in bytes on an mbarrier) for the NEXT block while the warp works through the
// current one, so the serial loop never waits for global memory and spends no instructions on the staging itself: a step is
// stage read -> hash -> table read -> match -> table write.
constexpr int LZC_STAGE_STEPS = 16;
constexpr uint32_t LZC_STAGE_BYTES = 32 * LZC_STAGE_STEPS + 32;   // 512 positions + 3 bytes, + up to 15 bytes of misalignment, in 16-byte units

__device__ __forceinline__ uint32_t lzc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lzc_mbar_init(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lzc_smem(bar)));
}
__device__ __forceinline__ void lzc_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {   // src, dst, bytes: multiples of 16
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lzc_smem(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(lzc_smem(dst)), "l"(src), "r"(bytes),
                 "r"(lzc_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void lzc_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LZC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LZC_DONE_%=;\n"
        "bra LZC_WAIT_%=;\n"
        "LZC_DONE_%=:\n"
        "}\n" ::"r"(lzc_smem(bar)),
        "r"(parity)
        : "memory");
}

__global__ void __launch_bounds__(32) lzc_hashlink_k(const uint8_t* __restrict__ bs, uint32_t n, const uint32_t* __restrict__ fs,
                                                     const LzcItem* __restrict__ items, uint32_t* __restrict__ lwh, uint16_t* __restrict__ rsd, int hb) {
    extern __shared__ uint32_t lzc_tab[];
    __shared__ __align__(16) uint8_t stage[2][LZC_STAGE_BYTES];
    __shared__ __align__(8) uint64_t bar[2];
    const LzcItem it = items[blockIdx.x];
    const uint32_t base = fs[it.frame], len = fs[it.frame + 1] - base;
    const uint32_t lane = threadIdx.x;
    for (uint32_t k = lane; k < (1u << hb); k += 32) lzc_tab[k] = 0;
    if (lane == 0) {
        lzc_mbar_init(&bar[0]);
        lzc_mbar_init(&bar[1]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t i0 = it.start > LZC_PREROLL ? it.start - LZC_PREROLL : 0u;
    uint32_t run_start = i0;
    // readable 16-byte units of the batch buffer: [lo16, hi16) (bs has at least 64 bytes of padding after its n bytes)
    const uintptr_t lo16 = reinterpret_cast<uintptr_t>(bs) & ~(uintptr_t)15;
    const uintptr_t hi16 = (reinterpret_cast<uintptr_t>(bs) + n + 48u) & ~(uintptr_t)15;
    // block of 16 steps starting at position j0: from the 16-byte unit that holds byte j0 - 1 (the byte before its first position)
    auto block_addr = [&](uint32_t j0) {
        const uintptr_t a = (reinterpret_cast<uintptr_t>(bs) + base + j0 - 1u) & ~(uintptr_t)15;
        return a < lo16 ? lo16 : a;
    };
    auto issue = [&](uint32_t j0, int b) {   // one lane
        const uintptr_t a = block_addr(j0);
        const uint32_t bytes = (uint32_t)min((uintptr_t)LZC_STAGE_BYTES, hi16 - a);
        lzc_bulk_load(stage[b], reinterpret_cast<const void*>(a), bytes, &bar[b]);
    };
    if (lane == 0) issue(i0, 0);
    uint32_t phase = 0;   // bit b: parity the next wait on bar[b] looks for
    int buf = 0;
    for (uint32_t j0 = i0; j0 < it.end; j0 += 32 * LZC_STAGE_STEPS, buf ^= 1) {
        // the other buffer was last read one block ago (every lane is past it: the loop below ends in __syncwarp); order those
        // generic-proxy reads before the async-proxy write, then start the next block's copy
        if (lane == 0 && j0 + 32 * LZC_STAGE_STEPS < it.end) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(j0 + 32 * LZC_STAGE_STEPS, buf ^ 1);
        }
        lzc_mbar_wait(&bar[buf], (phase >> buf) & 1u);
        phase ^= 1u << buf;
        const uint32_t* const sw = reinterpret_cast<const uint32_t*>(stage[buf]);
        // byte offset of position j0 - 1 inside the staged block (-1 only for the very first byte of the buffer)
        const int off0 = (int)(reinterpret_cast<uintptr_t>(bs) + base + j0 - 1u - block_addr(j0));
        const uint32_t jend = min(j0 + 32 * LZC_STAGE_STEPS, it.end);
        for (uint32_t s0 = j0; s0 < jend; s0 += 32) {
            const uint32_t i = s0 + lane;
            const int o = off0 + (int)(i - j0);
            // bytes i-1 .. i+2 (o == -1: the buffer's very first position, which has no byte before it)
            const uint32_t w = o < 0 ? sw[0] << 8 : __funnelshift_r(sw[o >> 2], sw[(o >> 2) + 1], (o & 3) * 8);
            const bool valid = i + 3u <= len && i < it.end;
            const uint32_t h = lzc_hash(w >> 8, hb);
            if (s0 + 32u <= it.start) {
                // pre-roll (it.start is a multiple of 32): only the table and the run start matter, and "most recent position
                // of this hash" is a maximum - one shared-memory atomic per lane, no vote, nothing for the next step to wait for
                if (valid) atomicMax(&lzc_tab[h], i + 1u);
                const unsigned balp = __ballot_sync(0xffffffffu, i < len && (i == 0 || ((w >> 8) & 0xFFu) != (w & 0xFFu)));
                if (balp) run_start = s0 + (31u - (uint32_t)__clz(balp));
                __syncwarp();
                continue;
            }
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
                lwh[base + i] = dist | (w >> 8 & 0xFFFFu) << 16;
                rsd[base + i] = (uint16_t)min(i - rs, 65535u);
            }
            __syncwarp();
        }
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
// - and most warps of a block - waiting for the slowest one. Both kernels are therefore persistent warps that take
// 512-position chunks from a global counter and work in two interleaved phases:
//   sweep: a coalesced pass over a chunk that takes every position's FIRST hop (four independent rounds in flight) and
//          finishes the positions it settles; the others are queued (shared memory, one word each);
//   drain: the queue is worked off with lane refill - a lane whose walk ends takes the next queued position at once, so
//          every hop instruction has (almost) all lanes doing useful work. When the queue runs low the warp sweeps its
//          next chunk while the walks still in progress simply keep their state.
constexpr int LZC_ROUNDS = 16;
constexpr int LZC_WCHUNK = 32 * LZC_ROUNDS;           // positions per chunk
constexpr int LZC_WARPS = LZC_THREADS / 32;
constexpr int LZC_BCHUNK = LZC_WCHUNK * LZC_WARPS;    // positions per block (grid sizing of the persistent kernels)
constexpr int LZC_MLP = 4;                            // rounds of a sweep in flight together
constexpr int LZC_QCAP = LZC_WCHUNK + 32;             // queue words per warp

// Op: uint32_t sweep(cbase, q) appends the chunk's unfinished positions to q and returns their number;
//     begin(p) loads a queued position's walk; step() takes one hop and returns true (after writing the result) when done.
template <int ROUNDS, class Op>
__device__ __forceinline__ void lzc_drive(Op& op, uint32_t n, uint32_t* __restrict__ counter, uint32_t* q, int refill_min = 1) {
    const uint32_t lane = lane_id();
    uint32_t qn = 0, qi = 0;
    bool more = true, busy = false;
    for (;;) {
        if (more && qn - qi < 32u) {   // keep the leftovers, sweep the next chunk
            const uint32_t left = qn - qi;
            const uint32_t keep = lane < left ? q[qi + lane] : 0u;
            __syncwarp();
            if (lane < left) q[lane] = keep;
            qn = left;
            qi = 0;
            uint32_t ch = 0;
            if (lane == 0) ch = atomicAdd(counter, 1u);
            ch = __shfl_sync(0xffffffffu, ch, 0);
            const uint64_t cb = (uint64_t)ch * (32 * ROUNDS);
            if (cb >= n) more = false;
            else qn += op.sweep((uint32_t)cb, q + qn);
            __syncwarp();
            continue;
        }
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        // refill in batches: a begin() costs the whole warp its instructions whether one lane or thirty-two take a position
        if (qi < qn && (__popc(idle) >= refill_min || idle == 0xffffffffu)) {
            const uint32_t my = qi + __popc(idle & lanemask_lt());
            if (!busy && my < qn) { op.begin(q[my]); busy = true; }
            qi = min(qn, qi + (uint32_t)__popc(idle));
        }
        if (!__any_sync(0xffffffffu, busy)) {
            if (!more) break;
            continue;
        }
        if (busy && op.step()) busy = false;
    }
}

template <int ROUNDS>
struct LzcLink3Op {
    const uint8_t* __restrict__ bs; const uint32_t* __restrict__ fs; uint32_t F, n;
    const uint32_t* __restrict__ lwh; const uint16_t* __restrict__ rsd; uint32_t* __restrict__ lw3;
    LzcLink3Walk wlk;
    uint32_t b23c;   // byte 2 | byte 3 << 8 | cap << 16 of the position being walked
    const uint32_t* __restrict__ cframe;   // frame of the first position of every LZC_WCHUNK-position chunk (lzc_cframe_k)
    __device__ __forceinline__ uint32_t cap_of(uint32_t f, uint32_t p) const {   // min(15, bytes left in p's frame); f = a frame at or before p's
        while (fs[f + 1] <= p) f++;
        return min(fs[f + 1] - p, (uint32_t)LZ_MAXLEN);
    }
    __device__ __forceinline__ void put(uint32_t p, uint32_t nd, uint32_t b2, uint32_t b3, uint32_t cap) const {
        lw3[p] = nd ? lzc_word(nd, b3, b2, cap) : lzc_dead(0u, 0u, b3);
    }
    __device__ __forceinline__ uint32_t sweep(uint32_t cbase, uint32_t* q) {
        const uint32_t lane = lane_id();
        const uint32_t f0 = cframe[cbase / LZC_WCHUNK];
        uint32_t qn = 0;
        for (int r0 = 0; r0 < ROUNDS; r0 += LZC_MLP) {
            uint32_t w[LZC_MLP], b23[LZC_MLP], cap[LZC_MLP], wk[LZC_MLP], k2[LZC_MLP];
#pragma unroll
            for (int j = 0; j < LZC_MLP; j++) {
                const uint32_t p = cbase + (r0 + j) * 32 + lane;
                w[j] = 0; b23[j] = 0; cap[j] = 0;
                if (p < n) { w[j] = lwh[p]; b23[j] = (uint32_t)bs[p + 2] | (uint32_t)bs[p + 3] << 8; cap[j] = cap_of(f0, p); }
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
                    if (!pend) put(p, nd, b23[j] & 0xFFu, b23[j] >> 8, cap[j]);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pend);
                if (pend) q[qn + __popc(bal & lanemask_lt())] = p;
                qn += __popc(bal);
            }
        }
        return qn;
    }
    __device__ __forceinline__ void begin(uint32_t p) {
        const uint32_t cp = cap_of(cframe[p / LZC_WCHUNK], p);
        b23c = (uint32_t)bs[p + 2] | (uint32_t)bs[p + 3] << 8 | cp << 16;
        wlk.start(p, lwh[p], b23c & 0xFFu, cp);
    }
    __device__ __forceinline__ bool step() {
        const int r = wlk.hop(bs, lwh, rsd);
        if (r == LZC_GO) return false;
        put(wlk.p, r == LZC_FOUND ? wlk.acc : 0u, b23c & 0xFFu, (b23c >> 8) & 0xFFu, b23c >> 16);
        return true;
    }
};

template <int ROUNDS>
struct LzcLevelOp {
    const uint8_t* __restrict__ bs; uint32_t n, L;
    const uint32_t* __restrict__ lw; const uint16_t* __restrict__ rsd; uint32_t* __restrict__ lw_next;
    const uint32_t* __restrict__ fs; const uint32_t* __restrict__ cframe;
    LzcLevelWalk wlk;
    uint32_t nbc;   // byte L+1 | cap << 8 of the position being walked
    // The sweep settles a position without a walk in two cases:
    //   * first hop: the most recent occurrence of its L-gram also carries its byte L (the link stays);
    //   * right neighbour (lzchain_core.h, "neighbour rule"): the level-L word of p+1 names the only candidate - its link if it
    //     is live (p's new link), its offset if it died at L-1 (then p dies at L with that offset) - and one byte compare
    //     (the byte before the candidate against byte p) confirms it.
    __device__ __forceinline__ uint32_t sweep(uint32_t cbase, uint32_t* q) {
        return cbase + LZC_WCHUNK + 16u <= n ? sweep_full(cbase, q) : sweep_tail(cbase, q);
    }
    // four bytes starting at byte address a (any alignment), through two aligned words
    static __device__ __forceinline__ uint32_t bytes4(const uint8_t* a) {
        const uintptr_t u = reinterpret_cast<uintptr_t>(a);
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(u & ~(uintptr_t)3);
        return __funnelshift_r(wp[0], wp[1], (uint32_t)(u & 3u) * 8u);
    }
    // a whole chunk well inside the buffer, in two steps. (1) Dense: every lane owns four consecutive positions of each
    // 128-position group, so the level words move as 16-byte loads / stores; a live position takes its first hop and, if that
    // does not settle it, is queued. (2) Compact: the queued positions, one per lane, try the right-neighbour rule (two more
    // words from L1, two byte gathers); what is left stays queued for the walks. The rule costs nothing for the majority
    // whose first hop succeeds.
    __device__ __forceinline__ uint32_t sweep_full(uint32_t cbase, uint32_t* q) {
        const uint32_t lane = lane_id();
        uint32_t qn = 0;
#pragma unroll 1
        for (uint32_t g0 = 0; g0 < (uint32_t)LZC_WCHUNK; g0 += 128) {
            const uint32_t p0 = cbase + g0 + lane * 4u;
            const uint4 wv = *reinterpret_cast<const uint4*>(lw + p0);
            const uint32_t nb4 = bytes4(bs + p0 + L + 1u);
            uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w}, wk[4], out[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t dist = lzc_link(w[i]);
                wk[i] = 0;
                if (dist) wk[i] = lw[p0 + i - dist];
            }
            unsigned pendm = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t dist = lzc_link(w[i]);
                const uint32_t c = (w[i] >> 16) & 0xFFu, cap = (w[i] >> 24) & 0xFu, nb = (nb4 >> (8 * i)) & 0xFFu;
                out[i] = lzc_carry(w[i], nb);   // final already: the result travels on (a queued position's word is rewritten later)
                if (dist) {
                    if (L + 1u <= cap && ((wk[i] >> 16) & 0xFFu) == c) out[i] = lzc_word(dist, nb, c, cap);
                    else pendm |= 1u << i;
                }
            }
            *reinterpret_cast<uint4*>(lw_next + p0) = make_uint4(out[0], out[1], out[2], out[3]);
            if (__any_sync(0xffffffffu, pendm != 0u)) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const bool pend = (pendm >> i) & 1u;
                    const unsigned bal = __ballot_sync(0xffffffffu, pend);
                    if (pend) q[qn + __popc(bal & lanemask_lt())] = p0 + i;
                    qn += __popc(bal);
                }
            }
        }
        __syncwarp();
        // (2) right-neighbour rule over the queue, compacting it in place (the write index never passes the read index)
        const uint32_t f0 = cframe[cbase / LZC_WCHUNK];   // frame of the chunk's first position
        uint32_t keep = 0;
        for (uint32_t b0 = 0; b0 < qn; b0 += 32) {
            const bool have = b0 + lane < qn;
            const uint32_t p = have ? q[b0 + lane] : 0u;
            bool pend = have;
            if (have) {
                const uint32_t w = lw[p], wr = lw[p + 1], xr = wr & 0xFFFFu;
                const uint32_t cap = (w >> 24) & 0xFu;
                const bool dead_r = wr & LZC_DEAD;
                const bool cand = dead_r ? (L >= 4u && ((wr >> 24) & 0xFu) == L - 1u) : (L + 1u <= cap);
                if (cand && xr && xr <= p && bs[p - xr] == bs[p]) {
                    uint32_t f = f0;
                    while (fs[f + 1] <= p) f++;
                    if (p - xr >= fs[f]) {   // the candidate lies inside p's frame
                        const uint32_t nb = bs[p + L + 1], c = (w >> 16) & 0xFFu;
                        lw_next[p] = dead_r ? lzc_dead(L, xr, nb) : lzc_word(xr, nb, c, cap);
                        pend = false;
                    }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pend);
            __syncwarp();
            if (pend) q[keep + __popc(bal & lanemask_lt())] = p;
            keep += __popc(bal);
            __syncwarp();
        }
        return keep;
    }
    __device__ __forceinline__ uint32_t sweep_tail(uint32_t cbase, uint32_t* q) {
        const uint32_t lane = lane_id();
        const uint32_t f0 = cframe[cbase / LZC_WCHUNK];   // frame of the chunk's first position
        uint32_t qn = 0;
        for (int r0 = 0; r0 < ROUNDS; r0 += LZC_MLP) {
            uint32_t w[LZC_MLP], nb[LZC_MLP], wk[LZC_MLP], x[LZC_MLP], bkq[LZC_MLP];
            uint32_t wnext = LZC_DEAD;   // word of the position after this group of rounds
            {
                const uint32_t pn = cbase + (r0 + LZC_MLP) * 32;
                if (pn < n) wnext = lw[pn];
            }
#pragma unroll
            for (int j = 0; j < LZC_MLP; j++) {
                const uint32_t p = cbase + (r0 + j) * 32 + lane;
                w[j] = LZC_DEAD; nb[j] = 0;
                if (p < n) { w[j] = lw[p]; nb[j] = bs[p + L + 1]; }
            }
#pragma unroll
            for (int j = 0; j < LZC_MLP; j++) {
                const uint32_t p = cbase + (r0 + j) * 32 + lane, dist = lzc_link(w[j]);
                wk[j] = 0;
                if (dist) wk[j] = lw[p - dist];
                // candidate from the right neighbour's word
                uint32_t wr = __shfl_down_sync(0xffffffffu, w[j], 1);
                const uint32_t w0 = j + 1 < LZC_MLP ? __shfl_sync(0xffffffffu, w[j + 1 < LZC_MLP ? j + 1 : j], 0) : wnext;
                if (lane == 31) wr = w0;
                x[j] = 0; bkq[j] = 1;
                if (dist) {
                    const uint32_t xr = wr & 0xFFFFu;
                    const bool cand = (wr & LZC_DEAD) ? (L >= 4u && ((wr >> 24) & 0xFu) == L - 1u) : (L + 1u <= ((w[j] >> 24) & 0xFu));
                    if (cand && xr && xr <= p) {
                        x[j] = xr | (wr & LZC_DEAD);
                        bkq[j] = (uint32_t)bs[p - xr] ^ (uint32_t)bs[p];   // 0: the byte before the candidate is p's byte
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < LZC_MLP; j++) {
                const uint32_t p = cbase + (r0 + j) * 32 + lane, dist = lzc_link(w[j]);
                const uint32_t c = (w[j] >> 16) & 0xFFu, cap = (w[j] >> 24) & 0xFu;
                bool pend = false;
                if (p < n) {
                    if (!dist) lw_next[p] = lzc_carry(w[j], nb[j]);   // final already: the result travels on
                    else if (L + 1u <= cap && ((wk[j] >> 16) & 0xFFu) == c) lw_next[p] = lzc_word(dist, nb[j], c, cap);
                    else {
                        pend = true;
                        if (x[j] && !bkq[j]) {
                            uint32_t f = f0;
                            while (fs[f + 1] <= p) f++;
                            const uint32_t xr = x[j] & 0xFFFFu;
                            if (p - xr >= fs[f]) {   // the candidate lies inside p's frame
                                lw_next[p] = (x[j] & LZC_DEAD) ? lzc_dead(L, xr, nb[j]) : lzc_word(xr, nb[j], c, cap);
                                pend = false;
                            }
                        }
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pend);
                if (pend) q[qn + __popc(bal & lanemask_lt())] = p;
                qn += __popc(bal);
            }
        }
        return qn;
    }
    __device__ __forceinline__ void begin(uint32_t p) {
        const uint32_t w = lw[p];
        nbc = (uint32_t)bs[p + L + 1] | ((w >> 24) & 0xFu) << 8;
        wlk.start(p, w, L);
    }
    __device__ __forceinline__ bool step() {
        const int r = wlk.hop(lw, rsd);
        if (r == LZC_GO) return false;
        lw_next[wlk.p] = r == LZC_FOUND ? lzc_word(wlk.acc, nbc & 0xFFu, wlk.c, nbc >> 8) : lzc_dead(L, wlk.last, nbc & 0xFFu);
        return true;
    }
};

template <int ROUNDS>
__global__ void __launch_bounds__(LZC_THREADS) lzc_link3_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs, uint32_t F, uint32_t n,
                                                           const uint32_t* __restrict__ lwh, const uint16_t* __restrict__ rsd,
                                                           uint32_t* __restrict__ lw3, const uint32_t* __restrict__ cframe, uint32_t* __restrict__ counter) {
    static_assert(32 * ROUNDS == LZC_WCHUNK, "cframe is indexed by LZC_WCHUNK-position chunks");
    __shared__ uint32_t q[LZC_WARPS][32 * ROUNDS + 32];
    LzcLink3Op<ROUNDS> op{bs, fs, F, n, lwh, rsd, lw3};
    op.cframe = cframe;
    lzc_drive<ROUNDS>(op, n, counter, q[threadIdx.x >> 5]);
}

template <int ROUNDS>
__global__ void __launch_bounds__(LZC_THREADS) lzc_level_k(const uint8_t* __restrict__ bs, uint32_t n, uint32_t L, const uint32_t* __restrict__ lw,
                                                           const uint16_t* __restrict__ rsd, uint32_t* __restrict__ lw_next, const uint32_t* __restrict__ fs,
                                                           const uint32_t* __restrict__ cframe, uint32_t* __restrict__ counter, int refill_min) {
    static_assert(32 * ROUNDS == LZC_WCHUNK, "cframe is indexed by LZC_WCHUNK-position chunks");
    __shared__ uint32_t q[LZC_WARPS][32 * ROUNDS + 32];
    LzcLevelOp<ROUNDS> op{bs, n, L, lw, rsd, lw_next, fs, cframe};
    lzc_drive<ROUNDS>(op, n, counter, q[threadIdx.x >> 5], refill_min);
}

// frame of the first position of every LZC_WCHUNK-position chunk
__global__ void lzc_cframe_k(const uint32_t* __restrict__ fs, uint32_t F, uint32_t n, uint32_t* __restrict__ cframe) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if ((uint64_t)c * LZC_WCHUNK < n) cframe[c] = lzc_frame_of(fs, F, c * LZC_WCHUNK);
}

__global__ void lzc_wbase_k(const uint32_t* __restrict__ fs, uint32_t F, uint32_t* __restrict__ wbase) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= F) wbase[i] = (uint32_t)(((uint64_t)fs[i] * 9u) >> 5) + 3u * i;
}

// token emission (bit writer: src/agmv_utils.c:86-112; token layout src/agmv_encode.c:146-165): lzss.cuh emits from the
// parse's last pass (lz_emit_mark_k) and walks the 15-byte matches' chains in lz_pack15_k
__device__ __forceinline__ void lzc_emit(uint32_t* __restrict__ out_words, uint32_t wb, uint32_t rel, uint32_t v, uint32_t nb) {
    const uint32_t w = wb + (rel >> 5), sh = rel & 31;
    atomicOr(&out_words[w], v << sh);
    if (sh + nb > 32) atomicOr(&out_words[w + 1], v >> (32 - sh));
}
// the parse's input: one byte per position, the final match length (level-15 words: dead -> its length, live -> 15)
__global__ void __launch_bounds__(256) lzc_bestlen_k(const uint32_t* __restrict__ lw15, uint32_t n, uint8_t* __restrict__ bestlen) {
    const uint32_t i = (blockIdx.x * 256u + threadIdx.x) * 4u;   // lw15 and bestlen are 16-byte aligned
    if (i >= n) return;
    const uint4 w = *reinterpret_cast<const uint4*>(lw15 + i);   // (the arrays are padded past n)
    auto len = [](uint32_t x) { return (x & LZC_DEAD) ? (x >> 24) & 0xFu : (uint32_t)LZ_MAXLEN; };
    *reinterpret_cast<uint32_t*>(bestlen + i) = len(w.x) | len(w.y) << 8 | len(w.z) << 16 | len(w.w) << 24;
}

}  // namespace agmvb
