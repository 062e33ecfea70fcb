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
#include "orbit.cuh"
#include "radix.cuh"
#include "scan.cuh"
#include <vector>
#include "lzchain.cuh"

namespace agmvb {

constexpr int LZ_LEVELS = 15;
constexpr uint32_t LZ_POS_MASK = 0x0FFFFFFFu;  // A[] words: position in the low 28 bits, min(15, bytes left in the frame) in the top 4
constexpr uint32_t LZ_MAX_BATCH = 1u << 27;
constexpr uint32_t LZ_RES = 1u << 27;          // match_rec: level << 28 | LZ_RES | offset = already resolved (lz_small_k); else level << 28 | group start
// work layout of the level kernels: 4096-element tiles (same tiling as radix.cuh / scan.cuh), 512 threads x 8 rounds
constexpr int LZ_THREADS = 512, LZ_WARPS = 16, LZ_ROUNDS = 8, LZ_WARP_SPAN = 256, LZ_TILE = 4096;
constexpr uint32_t LZ_ALIVE = 0x80000000u;     // GS / gs_tmp words: bit 31 = "a match of this level's length exists for the element"
constexpr uint32_t LZ_GS_MASK = 0x7FFFFFFFu;

struct LzChain {
    uint32_t* ticket;   // [1]
    uint32_t* st_max;   // [ntiles]       flag << 30 | head index
    uint32_t* st_cnt;   // [ntiles][256]  flag << 30 | count
};
struct LzWork {
    uint32_t cap_n = 0, cap_frames = 0;
    uint8_t* bestlen = nullptr;          // per position, filled after the last level (input of the parse)
    uint32_t* A[LZ_LEVELS + 1] = {};     // level arrays: positions grouped by their first L bytes
    uint32_t* GS[LZ_LEVELS + 1] = {};    // group start (index into A[L]) of every element of A[L]
    uint32_t* gs_tmp = nullptr;          // level-L group starts carried into level-(L+1) order
    uint32_t* gs_carry[2] = {};          // fused path: carried group words, ping-pong
    uint32_t* bstart = nullptr;          // fused path: [15][257] key-byte bucket starts per level
    uint32_t* bytehist = nullptr;        // fused path: [256]
    uint32_t* chain_mem = nullptr;       // fused path: look-back state (ticket, per-tile max, per-tile x 256 counts)
    bool fused = false;
    uint32_t* dig4[2] = {};              // next four key bytes of every element, carried through the scatters
    uint32_t* match_rec = nullptr;       // per position: best level << 28 | group start in A[best level]; 0 = no match
    uint32_t* bitcum = nullptr;          // per position: bit offset inside the frame's output if the parse visits it
    uint32_t* tile_hist[2] = {};         // key-byte counts / offsets per (digit, tile), double buffered across levels
    uint32_t* scan_ws = nullptr;
    uint32_t* run_ws = nullptr;          // run tables of the large three-equal-byte groups (cap_n words)
    bool runs = false;                   // AGMVB_LZ_RUNS: resolve those groups from run tables instead of the global levels
    bool legacy = false;                 // AGMVB_LZ_LEGACY=1: the radix-refinement match finder below instead of lzchain.cuh
    uint32_t* lw[2] = {};                // chain path: link words, ping-pong (lw[1] doubles as the hash-level links)
    uint16_t* rsd = nullptr;             // chain path: distance of every position to the start of its byte run
    LzcItem* items = nullptr;            // chain path: (frame, range) work items of lzc_hashlink_k
    uint32_t n_items = 0, cap_items = 0;
    uint32_t* counters = nullptr;        // chain path: chunk counters of the persistent kernels (one per launch of a batch)
    uint32_t link3_blocks = 148 * 5, level_blocks = 148 * 6;   // chain path: resident blocks of the persistent walking kernels
    OrbitTables orb;                     // greedy-parse tables (cap_n / ORB_TILE + cap_frames tiles)
    OrbitSeg* segs = nullptr;            // cap_frames
    uint32_t* seg_len = nullptr;         // cap_frames
    uint32_t* out_words = nullptr;
    size_t out_words_cap = 0;
    uint32_t* wbase = nullptr;           // cap_frames + 1
    uint32_t* outbits = nullptr;         // cap_frames
    uint32_t* csize = nullptr;           // cap_frames
    uint32_t* chunk_off = nullptr;       // cap_frames + 1
    uint32_t stub_bytes = 8;             // empty 'AGAC' chunk after every frame chunk (AGMV_EncodeAGMV); 0 for the other encoders
    uint32_t audio_chunk = 0;            // bytes of audio per frame (stub_bytes = 8 + audio_chunk while a track is interleaved)
    const uint8_t* audio = nullptr;      // companded track on the device (AGMV_CompressAudio's atsample)
    uint64_t audio_size = 0;
};

struct APtrs { const uint32_t* a[LZ_LEVELS + 1]; const uint32_t* gs[LZ_LEVELS + 1]; };

__device__ __forceinline__ uint32_t frame_of(const uint32_t* __restrict__ fs, uint32_t F, uint32_t i) {
    // largest f with fs[f] <= i and fs[f+1] > i (empty frames are skipped)
    uint32_t lo = 0, hi = F;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (fs[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t load4(const uint8_t* __restrict__ p) {
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

__global__ void lz_init_k(const uint8_t* __restrict__ bs, uint32_t n, const uint32_t* __restrict__ fs, uint32_t F, uint32_t* __restrict__ A0,
                          uint32_t* __restrict__ gs0, uint32_t* __restrict__ dig4, uint32_t* __restrict__ match_rec,
                          uint32_t* __restrict__ bitcum, uint32_t* __restrict__ wbase) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= F) wbase[i] = (uint32_t)(((uint64_t)fs[i] * 9u) >> 5) + 3u * i;
    if (i >= n) return;
    uint32_t f = frame_of(fs, F, i);
    uint32_t rem = fs[f + 1] - i;
    A0[i] = i | (rem < (uint32_t)LZ_MAXLEN ? rem : (uint32_t)LZ_MAXLEN) << 28;
    gs0[i] = fs[f];
    dig4[i] = load4(bs + i);
    match_rec[i] = 0;
    bitcum[i] = EMPTY32;
}

// key byte of level L: carried, no gather (used for the level-0 histogram only)
struct LzDigit {
    const uint32_t* dig4;
    uint32_t sh;  // 8 * (L & 3)
    __device__ uint32_t operator()(uint32_t i) const { return (dig4[i] >> sh) & 255u; }
};

// ---- one refinement level: stable scatter by the level's key byte --------------------------
// Same ranking scheme as radix_scatter_k (radix.cuh), specialised for the level arrays: the 16 key words of a
// thread are loaded up front (and reused as payload), positions / group starts are loaded eight at a time before
// any store so that the loads overlap, and every fourth level the next four key bytes are gathered from the
// bitstream (positions of a group are close to sorted, the 4-byte read stays inside one or two sectors).
// dynamic shared memory of lz_scatter_k: the tile's elements staged in bucket order so that the global writes of a
// bucket are contiguous (a warp's store covers consecutive addresses instead of up to 32 different buckets)
constexpr size_t LZ_SCATTER_SMEM = (size_t)LZ_TILE * 12 + LZ_TILE;

__global__ void __launch_bounds__(LZ_THREADS) lz_scatter_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ pos_in,
                                                           const uint32_t* __restrict__ gs_in, const uint32_t* __restrict__ dig_in,
                                                           uint32_t* __restrict__ pos_out, uint32_t* __restrict__ gs_out,
                                                           uint32_t* __restrict__ dig_out, uint32_t L, uint32_t n, uint32_t ntiles,
                                                           const uint32_t* __restrict__ tile_off) {
    extern __shared__ __align__(16) uint8_t lz_dyn[];
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(lz_dyn);
    uint32_t* s_gs = s_pos + LZ_TILE;
    uint32_t* s_dig = s_gs + LZ_TILE;
    uint8_t* s_d = reinterpret_cast<uint8_t*>(s_dig + LZ_TILE);
    __shared__ uint32_t wc[LZ_WARPS][256];
    __shared__ uint32_t goff[256];    // global index of the tile's first element of each bucket, minus its slot in the tile
    __shared__ uint32_t lstart[256];  // slot of the tile's first element of each bucket
    __shared__ uint32_t gtot[8];
    for (int k = threadIdx.x; k < LZ_WARPS * 256; k += LZ_THREADS) (&wc[0][0])[k] = 0;
    const int warp = threadIdx.x >> 5;
    const uint32_t t0 = blockIdx.x * LZ_TILE;
    const uint32_t base = t0 + warp * LZ_WARP_SPAN + lane_id();
    const uint32_t sh = 8u * (L & 3u);
    uint32_t word[LZ_ROUNDS], p[LZ_ROUNDS], g[LZ_ROUNDS];
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) {
        uint32_t i = base + r * 32;
        word[r] = i < n ? dig_in[i] : 0u;
        p[r] = i < n ? pos_in[i] : 0u;
        g[r] = i < n ? gs_in[i] : 0u;
    }
    __syncthreads();
    uint32_t packed[LZ_ROUNDS];  // digit << 16 | rank inside the warp's span
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) {
        uint32_t i = base + r * 32;
        bool valid = i < n;
        uint32_t key = valid ? (word[r] >> sh) & 255u : 256u + lane_id();
        unsigned peers = __match_any_sync(0xffffffffu, key);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && (int)lane_id() == leader) {
            old = wc[warp][key];
            wc[warp][key] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        packed[r] = (key << 16) | (old + __popc(peers & lanemask_lt()));
        __syncwarp();
    }
    __syncthreads();
    uint32_t b_tot = 0, b_inc = 0;  // per bucket (threads 0..255): tile total and its inclusive scan inside a group of 32 buckets
    if (threadIdx.x < 256) {
        const int d = threadIdx.x;
#pragma unroll
        for (int w = 0; w < LZ_WARPS; w++) {
            uint32_t t = wc[w][d];
            wc[w][d] = b_tot;
            b_tot += t;
        }
        b_inc = b_tot;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, b_inc, k); if ((int)lane_id() >= k) b_inc += y; }
        if (lane_id() == 31) gtot[d >> 5] = b_inc;
    }
    __syncthreads();
    if (threadIdx.x < 256) {  // slot of the bucket's first element inside the tile: exclusive scan over the 256 buckets
        uint32_t add = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) add += gtot[w];
        lstart[threadIdx.x] = add + b_inc - b_tot;
    }
    __syncthreads();
    if (threadIdx.x < 256) goff[threadIdx.x] = tile_off[threadIdx.x * ntiles + blockIdx.x] - lstart[threadIdx.x];
    const bool gather = (L & 3u) == 3u;
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) {
        uint32_t i = base + r * 32;
        if (i < n) {
            const uint32_t d = packed[r] >> 16, rk = packed[r] & 0xffffu;
            const uint32_t slot = lstart[d] + wc[warp][d] + rk;
            s_pos[slot] = p[r];
            s_gs[slot] = g[r];
            s_dig[slot] = gather ? load4(bs + (p[r] & LZ_POS_MASK) + L + 1) : word[r];
            s_d[slot] = (uint8_t)d;
        }
    }
    __syncthreads();
    const uint32_t cnt = min((uint32_t)LZ_TILE, n - t0);
    for (uint32_t k = threadIdx.x; k < cnt; k += LZ_THREADS) {
        const uint32_t dst = goff[s_d[k]] + k;
        pos_out[dst] = s_pos[k];
        gs_out[dst] = s_gs[k];
        dig_out[dst] = s_dig[k];
    }
}

// ---- group starts of the refined array ------------------------------------------
// After the scatter the array is ordered by (key byte, old group, position). An element heads a new group iff its
// old group differs from its predecessor's or it is the first element of a key-byte bucket (bucket starts come
// from the radix offsets: tile_off[d * ntiles]). New group start = running max over (head ? idx : 0): a device-wide
// inclusive max-scan (tile reduce, scan of the partials, tile apply), with the bucket-start bitmap in shared memory.
__device__ __forceinline__ void lz_bucket_bitmap(uint32_t* bm, const uint32_t* __restrict__ tile_off, uint32_t ntiles, uint32_t t0) {
    if (threadIdx.x < LZ_TILE / 32) bm[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < 256) {
        const uint32_t s = tile_off[threadIdx.x * ntiles];  // one bucket per thread
        if (s >= t0 && s < t0 + LZ_TILE) atomicOr(&bm[(s - t0) >> 5], 1u << ((s - t0) & 31));
    }
    __syncthreads();
}

// value of element (warp span, round r, this lane) and of its predecessor, from 16 up-front coalesced loads
__device__ __forceinline__ void lz_load_with_prev(const uint32_t* __restrict__ a, uint32_t base, uint32_t n, uint32_t cur[LZ_ROUNDS],
                                                  uint32_t prv[LZ_ROUNDS]) {
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) {
        uint32_t i = base + r * 32;
        cur[r] = i < n ? a[i] : 0u;
    }
    const uint32_t first = base - lane_id();  // index of lane 0's element in round 0
    uint32_t before = (lane_id() == 0 && first > 0 && first - 1 < n) ? a[first - 1] : 0u;
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) {
        uint32_t up = __shfl_up_sync(0xffffffffu, cur[r], 1);
        uint32_t wrap = r > 0 ? __shfl_sync(0xffffffffu, cur[r > 0 ? r - 1 : 0], 31) : before;
        prv[r] = lane_id() == 0 ? wrap : up;
    }
}

// head test shared by the reduce and apply kernels. BY_KEY (level 3 only, levels 1 and 2 run no group phase):
// old group = frame, refined by the 3-byte key carried in the key-byte word.
template <bool BY_KEY>
__device__ __forceinline__ bool lz_is_head(uint32_t gcur, uint32_t gprv, uint32_t kcur, uint32_t kprv, const uint32_t* bm, uint32_t t0, uint32_t idx) {
    if (((gcur ^ gprv) & LZ_GS_MASK) != 0) return true;
    if (BY_KEY) return ((kcur ^ kprv) & 0xFFFFFFu) != 0;
    return (bm[(idx - t0) >> 5] >> ((idx - t0) & 31)) & 1u;
}

template <bool BY_KEY>
__global__ void __launch_bounds__(LZ_THREADS) lz_group_reduce_k(const uint32_t* __restrict__ gs_old, const uint32_t* __restrict__ tile_off,
                                                                  uint32_t ntiles, uint32_t n, const uint32_t* __restrict__ dig,
                                                                  uint32_t* __restrict__ partial) {
    __shared__ uint32_t bm[LZ_TILE / 32];
    __shared__ uint32_t wmax[LZ_THREADS / 32];
    const uint32_t t0 = blockIdx.x * LZ_TILE;
    if (!BY_KEY) lz_bucket_bitmap(bm, tile_off, ntiles, t0);
    const int warp = threadIdx.x >> 5;
    const uint32_t base = t0 + warp * LZ_WARP_SPAN + lane_id();
    uint32_t cur[LZ_ROUNDS], prv[LZ_ROUNDS], kc[LZ_ROUNDS], kp[LZ_ROUNDS];
    lz_load_with_prev(gs_old, base, n, cur, prv);
    if (BY_KEY) lz_load_with_prev(dig, base, n, kc, kp);
    uint32_t acc = 0;
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) {
        uint32_t idx = base + r * 32;
        if (idx < n && idx > 0 && lz_is_head<BY_KEY>(cur[r], prv[r], BY_KEY ? kc[r] : 0u, BY_KEY ? kp[r] : 0u, bm, t0, idx))
            acc = idx;  // idx grows with r, so the last head seen is the maximum
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = max(acc, __shfl_xor_sync(0xffffffffu, acc, d));
    if (lane_id() == 0) wmax[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < LZ_THREADS / 32; w++) t = max(t, wmax[w]);
        partial[blockIdx.x] = t;
    }
}

// Writes GS[Lnew] (group start | "a match of length Lnew exists" in bit 31). exists_L is monotone in L, so the
// longest match of a position is the last level at which it exists: it is recorded exactly once, at the level
// where it stops existing (with the previous level's group start, which the scatter carried along) or at level 15.
// Also counts the next level's key bytes of this tile (the tile is exactly the next scatter's tile).
template <bool BY_KEY, bool NEXT_HIST>
__global__ void __launch_bounds__(LZ_THREADS) lz_group_apply_k(const uint32_t* __restrict__ gs_old, const uint32_t* __restrict__ tile_off,
                                                                 uint32_t ntiles, uint32_t n, const uint32_t* __restrict__ partial,
                                                                 const uint32_t* __restrict__ pos, uint32_t* __restrict__ gs_new,
                                                                 uint32_t* __restrict__ match_rec, uint32_t Lnew,
                                                                 const uint32_t* __restrict__ dig_next, uint32_t* __restrict__ hist_next) {
    __shared__ uint32_t bm[LZ_TILE / 32];
    __shared__ uint32_t wtot[LZ_THREADS / 32];
    __shared__ uint32_t h[256];
    const uint32_t t0 = blockIdx.x * LZ_TILE;
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    if (!BY_KEY) lz_bucket_bitmap(bm, tile_off, ntiles, t0); else __syncthreads();
    const int warp = threadIdx.x >> 5;
    const uint32_t base = t0 + warp * LZ_WARP_SPAN + lane_id();
    uint32_t cur[LZ_ROUNDS], prv[LZ_ROUNDS], kc[LZ_ROUNDS], kp[LZ_ROUNDS];
    lz_load_with_prev(gs_old, base, n, cur, prv);
    if (BY_KEY || NEXT_HIST) {
#pragma unroll
        for (int r = 0; r < LZ_ROUNDS; r++) { uint32_t i = base + r * 32; kc[r] = i < n ? dig_next[i] : 0u; }
    }
    if (BY_KEY) {
        const uint32_t first = base - lane_id();
        uint32_t before = (lane_id() == 0 && first > 0 && first - 1 < n) ? dig_next[first - 1] : 0u;
#pragma unroll
        for (int r = 0; r < LZ_ROUNDS; r++) {
            uint32_t up = __shfl_up_sync(0xffffffffu, kc[r], 1);
            uint32_t wrap = r > 0 ? __shfl_sync(0xffffffffu, kc[r > 0 ? r - 1 : 0], 31) : before;
            kp[r] = lane_id() == 0 ? wrap : up;
        }
    }
    uint32_t v[LZ_ROUNDS];
    uint32_t alive_old = 0;  // bit r: the element existed at level Lnew-1
    uint32_t carry = 0;
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) {
        uint32_t idx = base + r * 32;
        const bool hd = idx < n && idx > 0 && lz_is_head<BY_KEY>(cur[r], prv[r], BY_KEY ? kc[r] : 0u, BY_KEY ? kp[r] : 0u, bm, t0, idx);
        alive_old |= (cur[r] >> 31) << r;
        // running max of (head ? idx : 0) inside the warp span: the last head at or before this lane, else the carry
        const unsigned hm = __ballot_sync(0xffffffffu, hd);
        const unsigned upto = hm & (0xffffffffu >> (31 - lane_id()));
        const uint32_t rbase = idx - lane_id();
        v[r] = upto ? rbase + (31 - __clz(upto)) : carry;
        carry = hm ? rbase + (31 - __clz(hm)) : carry;
    }
    if (lane_id() == 0) wtot[warp] = carry;
    uint32_t lo_old[LZ_ROUNDS];
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) lo_old[r] = cur[r] & LZ_GS_MASK;
    lz_load_with_prev(pos, base, n, cur, prv);  // reuse the registers: positions and predecessor positions
    __syncthreads();
    uint32_t pre = partial[blockIdx.x];
    for (int w = 0; w < warp; w++) pre = max(pre, wtot[w]);
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) {
        uint32_t idx = base + r * 32;
        if (idx < n) {
            const uint32_t g = max(pre, v[r]);
            const uint32_t p = cur[r] & LZ_POS_MASK, prev = prv[r] & LZ_POS_MASK;
            const bool exists = g != idx && p - prev <= (uint32_t)LZ_WINDOW && Lnew <= (cur[r] >> 28);
            gs_new[idx] = g | (exists ? LZ_ALIVE : 0u);
            if (!exists && ((alive_old >> r) & 1u)) match_rec[p] = (Lnew - 1) << 28 | lo_old[r];
            if (exists && Lnew == (uint32_t)LZ_LEVELS) match_rec[p] = Lnew << 28 | g;
        }
    }
    if (NEXT_HIST) {
        const uint32_t sh = 8u * (Lnew & 3u);
#pragma unroll
        for (int r = 0; r < LZ_ROUNDS; r++) {
            uint32_t idx = base + r * 32;
            bool valid = idx < n;
            uint32_t key = valid ? (kc[r] >> sh) & 255u : 256u + lane_id();
            unsigned peers = __match_any_sync(0xffffffffu, key);
            if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&h[key], __popc(peers));
        }
        __syncthreads();
        if (threadIdx.x < 256) hist_next[threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
    }
}

__global__ void lz_bestlen_k(const uint32_t* __restrict__ match_rec, uint32_t n, uint8_t* __restrict__ bestlen) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bestlen[i] = (uint8_t)(match_rec[i] >> 28);
}

// =====================================================================================================
// Levels 4..15 for SMALL groups, entirely in shared memory.
//
// After level 3 the array is cut into groups (same frame, same first 3 bytes); every later level only splits groups,
// so a group is an independent sub-problem. Groups of at most SG_C elements - 75-100 % of all positions on the bench
// workload - are finished by one CTA without touching HBM again: CTA w owns the groups that START in
// [w*SG_W, (w+1)*SG_W) (they end before (w+1)*SG_W + SG_C, so SG_W + SG_C slots hold them), compacts them into shared memory and
// runs the same refinement as the global kernels (stable counting sort by the next key byte, heads = old-group change
// or bucket start, running max, exists = predecessor within the window) twelve times. When a position's match stops
// growing it is resolved on the spot: the earliest in-window member of its last group is a binary search in the
// previous arrangement (still in shared memory), so match_rec receives the final (length, offset) and lz_pack_k has
// nothing to search. Larger groups (long runs, mostly) are compacted and go through the global level kernels.
// =====================================================================================================
constexpr int SG_C = 512;                // largest group finished in shared memory
constexpr int SG_W = 1536;               // window of group starts per CTA
constexpr int SG_K = SG_W + SG_C;        // slots per CTA
constexpr int SG_THREADS = 256, SG_WARPS = SG_THREADS / 32, SG_ROUNDS = SG_K / SG_THREADS, SG_SPAN = SG_ROUNDS * 32;
constexpr size_t SG_SMEM = (size_t)SG_K * (4 + 4) * 2 + (size_t)SG_K * 2 * 2 + (size_t)SG_K * 2;  // pos, kw (x2), gid (x2), src
static_assert(SG_K == SG_THREADS * SG_ROUNDS && SG_K <= 4096, "one slot per thread and round; local indices fit 12 bits");

__device__ __forceinline__ uint32_t sg_lower_bound(const uint32_t* __restrict__ pos, uint32_t lo, uint32_t hi, uint32_t target) {
    // first k in [lo, hi) with (pos[k] & LZ_POS_MASK) >= target; the caller guarantees that one exists
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((pos[mid] & LZ_POS_MASK) >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// is the level-3 group that starts at index g larger than SG_C?
__device__ __forceinline__ bool sg_group_is_large(const uint32_t* __restrict__ gs3, uint32_t n, uint32_t g) {
    return g + (uint32_t)SG_C < n && (gs3[g + SG_C] & LZ_GS_MASK) == g;
}
struct LargeFlag {
    const uint32_t* gs3;
    uint32_t n;
    __device__ uint32_t operator()(uint32_t i) const { return sg_group_is_large(gs3, n, gs3[i] & LZ_GS_MASK) ? 1u : 0u; }
};
// elements of large groups, in order, with group starts renumbered; prefix = exclusive scan of LargeFlag
__global__ void __launch_bounds__(256) lz_compact_large_k(const uint32_t* __restrict__ a3, const uint32_t* __restrict__ gs3,
                                                          const uint32_t* __restrict__ dig, const uint32_t* __restrict__ prefix, uint32_t n,
                                                          uint32_t* __restrict__ a_out, uint32_t* __restrict__ gs_out, uint32_t* __restrict__ dig_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t gw = gs3[i], g = gw & LZ_GS_MASK;
    if (!sg_group_is_large(gs3, n, g)) return;
    const uint32_t o = prefix[i];
    a_out[o] = a3[i];
    gs_out[o] = prefix[g] | (gw & LZ_ALIVE);
    dig_out[o] = dig[i];
}

// ---- tiny groups: all pairs ---------------------------------------------------------------------------
// A level-3 group of at most SG_TINY positions needs no refinement at all: a position's longest match is the largest
// common prefix with any earlier member inside the window, a handful of comparisons. One thread per position walks its
// group backwards (most recent member first) while the members are within 65535 bytes; ">=" on the length leaves the
// EARLIEST start among the longest, as the reference's ascending scan with ">" does.
constexpr int SG_TINY = 128;
__device__ __forceinline__ bool sg_group_is_tiny(const uint32_t* __restrict__ gs3, uint32_t n, uint32_t g) {
    return g + (uint32_t)SG_TINY >= n || (gs3[g + SG_TINY] & LZ_GS_MASK) != g;
}
__device__ __forceinline__ void sg_key12(const uint8_t* __restrict__ p, uint32_t k[3]) {  // bytes 3..14 after position p
    const uintptr_t a = reinterpret_cast<uintptr_t>(p + 3);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    k[0] = __funnelshift_r(w0, w1, sh);
    k[1] = __funnelshift_r(w1, w2, sh);
    k[2] = __funnelshift_r(w2, w3, sh);
}
__global__ void __launch_bounds__(256) lz_tiny_k(const uint8_t* __restrict__ bs, uint32_t n, const uint32_t* __restrict__ a3,
                                                 const uint32_t* __restrict__ gs3, uint32_t* __restrict__ match_rec) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const uint32_t gw = gs3[idx];
    if (!(gw & LZ_ALIVE)) return;                       // no earlier member within the window: no match of length 3
    const uint32_t g = gw & LZ_GS_MASK;
    if (!sg_group_is_tiny(gs3, n, g)) return;
    const uint32_t pw = a3[idx], p = pw & LZ_POS_MASK, cap = pw >> 28;   // cap = min(15, bytes left in the frame) >= 3 here
    uint32_t ky[3];
    sg_key12(bs + p, ky);
    uint32_t best = 0, bestx = 0;
    for (uint32_t k = idx; k-- > g;) {
        const uint32_t x = a3[k] & LZ_POS_MASK;
        if (p - x > (uint32_t)LZ_WINDOW) break;
        uint32_t kx[3];
        sg_key12(bs + x, kx);
        uint32_t l = 12;
        const uint32_t d2 = kx[2] ^ ky[2], d1 = kx[1] ^ ky[1], d0 = kx[0] ^ ky[0];
        if (d2) l = 8 + ((uint32_t)(__ffs((int)d2) - 1) >> 3);
        if (d1) l = 4 + ((uint32_t)(__ffs((int)d1) - 1) >> 3);
        if (d0) l = (uint32_t)(__ffs((int)d0) - 1) >> 3;
        l = min(l + 3u, cap);
        if (l >= best) { best = l; bestx = x; }
    }
    match_rec[p] = best << 28 | LZ_RES | (p - bestx);
}

__global__ void __launch_bounds__(SG_THREADS, 4) lz_small_k(const uint8_t* __restrict__ bs, uint32_t n, const uint32_t* __restrict__ a3,
                                                         const uint32_t* __restrict__ gs3, uint32_t* __restrict__ match_rec) {
    extern __shared__ __align__(16) uint8_t lz_dyn[];
    uint32_t* const pos_all = reinterpret_cast<uint32_t*>(lz_dyn);            // [2][SG_K]
    uint32_t* const kw_all = pos_all + 2 * SG_K;                              // [2][SG_K]
    uint16_t* const gid_all = reinterpret_cast<uint16_t*>(kw_all + 2 * SG_K);  // [2][SG_K]
    uint16_t* const src = gid_all + 2 * SG_K;                                  // [SG_K]
    __shared__ uint32_t n_list;
    __shared__ uint32_t wc[SG_WARPS][256];
    __shared__ uint32_t lstart[256];
    __shared__ uint32_t gtot[8];
    __shared__ uint32_t bm[SG_K / 32];
    __shared__ uint32_t wtot[SG_WARPS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w0 = blockIdx.x * (uint32_t)SG_W;
    const uint32_t sbase = (uint32_t)warp * SG_SPAN + lane;   // slot of round r: sbase + 32 r

    // ---- load the window, keep the small groups that start in it ----
    uint32_t pw[SG_ROUNDS], gw[SG_ROUNDS];
#pragma unroll
    for (int r = 0; r < SG_ROUNDS; r++) {
        const uint32_t s = sbase + r * 32, idx = w0 + s;
        gw[r] = idx < n ? gs3[idx] : 0xFFFFFFFFu;
        pw[r] = idx < n ? a3[idx] : 0u;
        pos_all[SG_K + s] = gw[r];
    }
    __syncthreads();
    uint32_t keep = 0;  // bit r
    uint32_t lidx[SG_ROUNDS];
    {
        uint32_t run = 0;
#pragma unroll
        for (int r = 0; r < SG_ROUNDS; r++) {
            const uint32_t s = sbase + r * 32, idx = w0 + s, g = gw[r] & LZ_GS_MASK;
            bool v = idx < n && g >= w0 && g < w0 + (uint32_t)SG_W;
            if (v) v = g + (uint32_t)SG_C >= n || (pos_all[SG_K + g + SG_C - w0] & LZ_GS_MASK) != g;          // not large
            if (v) v = g + (uint32_t)SG_TINY < n && (pos_all[SG_K + g + SG_TINY - w0] & LZ_GS_MASK) == g;      // not tiny (lz_tiny_k)
            const unsigned bal = __ballot_sync(0xffffffffu, v);
            lidx[r] = run + __popc(bal & lanemask_lt());
            run += __popc(bal);
            keep |= (uint32_t)v << r;
        }
        if (lane == 0) wtot[warp] = run;
    }
    __syncthreads();
    uint32_t wpre = 0, K = 0;
#pragma unroll
    for (int w = 0; w < SG_WARPS; w++) { const uint32_t t = wtot[w]; if (w < warp) wpre += t; K += t; }
    if (K == 0) return;
#pragma unroll
    for (int r = 0; r < SG_ROUNDS; r++) {
        lidx[r] += wpre;
        if ((keep >> r) & 1u) src[sbase + r * 32] = (uint16_t)lidx[r];   // slot -> local index (read below for the group heads)
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SG_ROUNDS; r++) {
        if ((keep >> r) & 1u) {
            const uint32_t i = lidx[r], g = gw[r] & LZ_GS_MASK;
            pos_all[i] = pw[r];
            gid_all[i] = (uint16_t)(src[g - w0] | (gw[r] >> 31) << 15);
            kw_all[i] = load4(bs + (pw[r] & LZ_POS_MASK) + 3);
        }
    }
    __syncthreads();

    // The elements are dealt out in equal contiguous spans (R rounds of 32 per warp). An element that has no match at the
    // current length and whose successor in its group is farther than the window (or missing) can never be anyone's
    // match source again - every later member of any sub-group is at least as far away - so it leaves the working set
    // at the next sort: K shrinks from level to level.
    __shared__ uint32_t s_knew;
    int cur = 0;
    for (uint32_t lvl = 3; lvl < (uint32_t)LZ_LEVELS; lvl++) {   // level lvl -> lvl + 1, key byte index lvl
        uint32_t* const posc = pos_all + cur * SG_K;  uint32_t* const posn = pos_all + (cur ^ 1) * SG_K;
        uint32_t* const kwc = kw_all + cur * SG_K;    uint32_t* const kwn = kw_all + (cur ^ 1) * SG_K;
        uint16_t* const gidc = gid_all + cur * SG_K;  uint16_t* const gidn = gid_all + (cur ^ 1) * SG_K;
        const int R = (int)((K + SG_THREADS - 1) / SG_THREADS);
        const uint32_t jbase = (uint32_t)warp * (uint32_t)(R * 32) + lane;
        if (tid == 0) n_list = 0;
        for (int k = tid; k < SG_WARPS * 256; k += SG_THREADS) (&wc[0][0])[k] = 0;
        if (tid < SG_K / 32) bm[tid] = 0;
        const bool regather = lvl > 3 && ((lvl - 3) & 3u) == 0;
        const uint32_t sh = 8u * ((lvl - 3) & 3u);
        uint32_t p[SG_ROUNDS], kk[SG_ROUNDS], packed[SG_ROUNDS];
        uint16_t gg[SG_ROUNDS];
        uint32_t stay = 0;  // bit r: the element takes part in the next arrangement
#pragma unroll
        for (int r = 0; r < SG_ROUNDS; r++) {
            if (r >= R) break;
            const uint32_t j = jbase + r * 32;
            if (j < K) {
                p[r] = posc[j];
                gg[r] = gidc[j];
                bool st = gg[r] >> 15;
                if (!st && j + 1 < K) {
                    const uint32_t gn = gidc[j + 1], pn = posc[j + 1];
                    st = ((gn ^ gg[r]) & 0x7FFFu) == 0 && (pn & LZ_POS_MASK) - (p[r] & LZ_POS_MASK) <= (uint32_t)LZ_WINDOW;
                }
                stay |= (uint32_t)st << r;
                kk[r] = !st ? 0u : (regather ? load4(bs + (p[r] & LZ_POS_MASK) + lvl) : kwc[j]);
            } else { p[r] = 0; gg[r] = 0; kk[r] = 0; }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SG_ROUNDS; r++) {
            if (r >= R) break;
            const bool valid = (stay >> r) & 1u;
            const uint32_t key = valid ? (kk[r] >> sh) & 255u : 256u + lane;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = wc[warp][key];
                wc[warp][key] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            packed[r] = (key << 16) | (old + __popc(peers & lanemask_lt()));
            __syncwarp();
        }
        __syncthreads();
        uint32_t b_tot = 0, b_inc = 0;
        if (tid < 256) {
#pragma unroll
            for (int w = 0; w < SG_WARPS; w++) {
                const uint32_t t = wc[w][tid];
                wc[w][tid] = b_tot;
                b_tot += t;
            }
            b_inc = b_tot;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, b_inc, k); if (lane >= k) b_inc += y; }
            if (lane == 31) gtot[tid >> 5] = b_inc;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t add = 0;
            for (int w = 0; w < (tid >> 5); w++) add += gtot[w];
            const uint32_t ls = add + b_inc - b_tot;
            lstart[tid] = ls;
            if (b_tot) atomicOr(&bm[ls >> 5], 1u << (ls & 31));   // the first element of a key-byte bucket heads a group
            if (tid == 255) s_knew = ls + b_tot;
        }
        __syncthreads();
        const uint32_t Kn = s_knew;
        if (Kn == 0) break;
#pragma unroll
        for (int r = 0; r < SG_ROUNDS; r++) {
            if (r >= R) break;
            if ((stay >> r) & 1u) {
                const uint32_t d = packed[r] >> 16, rk = packed[r] & 0xffffu;
                const uint32_t slot = lstart[d] + wc[warp][d] + rk;
                posn[slot] = p[r];
                kwn[slot] = kk[r];
                gidn[slot] = gg[r];
                src[slot] = (uint16_t)(jbase + r * 32);
            }
        }
        __syncthreads();
        // ---- groups of the new arrangement (Kn elements) ----
        const int Rn = (int)((Kn + SG_THREADS - 1) / SG_THREADS);
        const uint32_t nbase = (uint32_t)warp * (uint32_t)(Rn * 32) + lane;
        uint32_t v[SG_ROUNDS], pprev[SG_ROUNDS];
        {
            uint32_t carry = 0;
#pragma unroll
            for (int r = 0; r < SG_ROUNDS; r++) {
                if (r >= Rn) break;
                const uint32_t j = nbase + r * 32;
                const bool valid = j < Kn;
                p[r] = valid ? posn[j] : 0u;
                gg[r] = valid ? gidn[j] : (uint16_t)0;
                const uint32_t gp = (valid && j > 0) ? gidn[j - 1] : 0u;
                pprev[r] = (valid && j > 0) ? posn[j - 1] : 0u;
                const bool hd = valid && (j == 0 || ((gg[r] ^ gp) & 0x7FFFu) != 0 || ((bm[j >> 5] >> (j & 31)) & 1u));
                const unsigned hm = __ballot_sync(0xffffffffu, hd);
                const unsigned upto = hm & (0xffffffffu >> (31 - lane));
                const uint32_t rbase = j - lane;
                v[r] = upto ? rbase + (31 - __clz(upto)) : carry;
                carry = hm ? rbase + (31 - __clz(hm)) : carry;
            }
            if (lane == 0) wtot[warp] = carry;
        }
        __syncthreads();
        uint32_t pre = 0;
        for (int w = 0; w < warp; w++) pre = max(pre, wtot[w]);
        const uint32_t Lnew = lvl + 1;
        int any_alive = 0;
        uint16_t ng[SG_ROUNDS];
#pragma unroll
        for (int r = 0; r < SG_ROUNDS; r++) {
            if (r >= Rn) break;
            const uint32_t j = nbase + r * 32;
            ng[r] = 0;
            bool died = false, full = false;
            uint32_t g = 0;
            if (j < Kn) {
                g = max(pre, v[r]);
                const uint32_t pp = p[r] & LZ_POS_MASK;
                const bool exists = g != j && pp - (pprev[r] & LZ_POS_MASK) <= (uint32_t)LZ_WINDOW && Lnew <= (p[r] >> 28);
                died = !exists && (gg[r] >> 15);
                full = exists && Lnew == (uint32_t)LZ_LEVELS;
                ng[r] = (uint16_t)(g | (exists ? 0x8000u : 0u));
                any_alive |= exists;
            }
            // a match that stops growing here (or reaches 15) is resolved after the loop, densely: the searches of the few
            // lanes that need one would otherwise stall the whole warp at every level
            const unsigned act = __ballot_sync(0xffffffffu, died || full);
            if (act) {
                const int leader = __ffs(act) - 1;
                uint32_t base_slot = 0;
                if (lane == leader) base_slot = atomicAdd(&n_list, (uint32_t)__popc(act));
                base_slot = __shfl_sync(0xffffffffu, base_slot, leader);
                if (died || full) kwc[base_slot + __popc(act & lanemask_lt())] = j | g << 12 | (full ? 1u << 24 : 0u);   // kwc is free after the scatter
            }
        }
        __syncthreads();
        for (uint32_t e = tid; e < n_list; e += SG_THREADS) {
            const uint32_t rec = kwc[e], j = rec & 0xFFFu, g = (rec >> 12) & 0xFFFu;
            const uint32_t pp = posn[j] & LZ_POS_MASK;
            const uint32_t target = pp > (uint32_t)LZ_WINDOW ? pp - (uint32_t)LZ_WINDOW : 0u;
            if (rec >> 24) {   // 15 bytes match: earliest in-window member of the level-15 group (this arrangement)
                const uint32_t k = sg_lower_bound(posn, g, j, target);
                match_rec[pp] = Lnew << 28 | LZ_RES | (pp - (posn[k] & LZ_POS_MASK));
            } else {           // the match stops at length lvl: earliest in-window member of the level-lvl group (previous arrangement)
                const uint32_t i = src[j], go = gidn[j] & 0x7FFFu;
                const uint32_t k = sg_lower_bound(posc, go, i, target);
                match_rec[pp] = lvl << 28 | LZ_RES | (pp - (posc[k] & LZ_POS_MASK));
            }
        }
        __syncthreads();   // every neighbour's old group id has been read
#pragma unroll
        for (int r = 0; r < SG_ROUNDS; r++) {
            if (r >= Rn) break;
            const uint32_t j = nbase + r * 32;
            if (j < Kn) gidn[j] = ng[r];
        }
        if (!__syncthreads_or(any_alive)) break;   // nobody can grow any further
        K = Kn;
        cur ^= 1;
    }
}

// =====================================================================================================
// Fused level kernel: group phase of level `lvl` + stable scatter by key byte `lvl`, one pass over the data.
//
// The three-kernel formulation above (scatter / group reduce / group apply, plus offset scans) reads and writes
// every element several times per level. Both cross-tile dependencies - the running max of head indices and the
// per-key-byte offsets of the scatter - are prefix computations over tiles, so they can be resolved with a
// decoupled look-back (tiles take tickets in order; each publishes its aggregate, then folds its predecessors'
// until it meets an inclusive prefix). The global bucket bases need no pass at all: the number of elements whose
// key byte `lvl` equals d is the byte histogram of bs[lvl .. n+lvl), known before the first level.
// Per element and level: 12 B read, 4 B (GS) + 12 B (scattered) written.
// =====================================================================================================
constexpr uint32_t LB_AGG = 1u << 30, LB_INCL = 2u << 30, LB_VAL = (1u << 30) - 1u;

__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_vol(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

// byte histogram of bs[0 .. n+15) -> base[256]; then per level the counts of bs[lvl .. n+lvl) and their exclusive scan
__global__ void __launch_bounds__(256) lz_bytehist_k(const uint8_t* __restrict__ bs, uint32_t total, uint32_t* __restrict__ base) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (total + 3) / 4; i += gridDim.x * blockDim.x) {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) if (i * 4 + j < total) w |= (uint32_t)bs[i * 4 + j] << (8 * j);
        const uint32_t cnt = min(4u, total - i * 4);
        for (uint32_t j = 0; j < cnt; j++) atomicAdd(&h[(w >> (8 * j)) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&base[threadIdx.x], h[threadIdx.x]);
}
// bstart[lvl][d] (d = 0..256): first index of key byte d in the order sorted by byte offset lvl
__global__ void __launch_bounds__(256) lz_bstart_k(const uint8_t* __restrict__ bs, uint32_t n, const uint32_t* __restrict__ base,
                                                   uint32_t* __restrict__ bstart) {
    __shared__ uint32_t c[256];
    const uint32_t d = threadIdx.x, lvl = blockIdx.x;
    uint32_t v = base[d];
    for (uint32_t j = 0; j < lvl; j++) v -= bs[j] == d;                      // bytes before the range
    for (uint32_t j = n + lvl; j < n + LZ_LEVELS; j++) v -= bs[j] == d;      // bytes after it
    c[d] = v;
    __syncthreads();
    if (d == 0) {
        uint32_t run = 0;
        for (int k = 0; k < 256; k++) { uint32_t t = c[k]; c[k] = run; run += t; }
        bstart[lvl * 257 + 256] = run;
    }
    __syncthreads();
    bstart[lvl * 257 + d] = c[d];
}

template <bool GROUP, bool BY_KEY, bool SCATTER>
__global__ void __launch_bounds__(LZ_THREADS) lz_level_k(const uint8_t* __restrict__ bs, uint32_t n, uint32_t lvl,
                                                         const uint32_t* __restrict__ pos_in, const uint32_t* __restrict__ gs_in,
                                                         const uint32_t* __restrict__ dig_in, uint32_t* __restrict__ gs_final,
                                                         uint32_t* __restrict__ match_rec, uint32_t* __restrict__ pos_out,
                                                         uint32_t* __restrict__ gs_out, uint32_t* __restrict__ dig_out,
                                                         const uint32_t* __restrict__ bstart, LzChain ch) {
    __shared__ uint32_t wc[LZ_WARPS][256];
    __shared__ uint32_t goff[256];
    __shared__ uint32_t bm[LZ_TILE / 32];
    __shared__ uint32_t wtot[LZ_WARPS];
    __shared__ uint32_t s_tile, s_pre;
    if (threadIdx.x == 0) s_tile = atomicAdd(ch.ticket, 1u);
    if (SCATTER) for (int k = threadIdx.x; k < LZ_WARPS * 256; k += LZ_THREADS) (&wc[0][0])[k] = 0;
    if (GROUP && !BY_KEY && threadIdx.x < LZ_TILE / 32) bm[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = s_tile, t0 = tile * LZ_TILE;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t base = t0 + warp * LZ_WARP_SPAN + lane;
    if (GROUP && !BY_KEY && threadIdx.x < 256) {   // heads at the key-byte bucket starts of the CURRENT order (sorted by byte lvl-1)
        const uint32_t s = bstart[(lvl - 1) * 257 + threadIdx.x];
        if (s >= t0 && s < t0 + LZ_TILE) atomicOr(&bm[(s - t0) >> 5], 1u << ((s - t0) & 31));
    }
    uint32_t pw[LZ_ROUNDS], gw[LZ_ROUNDS], dw[LZ_ROUNDS];
#pragma unroll
    for (int r = 0; r < LZ_ROUNDS; r++) {
        const uint32_t i = base + r * 32;
        pw[r] = i < n ? pos_in[i] : 0u;
        gw[r] = i < n ? gs_in[i] : 0u;
        dw[r] = i < n ? dig_in[i] : 0u;
    }
    uint32_t v[LZ_ROUNDS];      // GROUP: running max of head indices inside the warp span
    uint32_t pprev[LZ_ROUNDS];  // GROUP: predecessor position words
    if (GROUP) {
        const uint32_t first = base - lane;
        uint32_t gb = 0, pb = 0, db = 0;
        if (lane == 0 && first > 0 && first - 1 < n) { gb = gs_in[first - 1]; pb = pos_in[first - 1]; if (BY_KEY) db = dig_in[first - 1]; }
        __syncthreads();  // bitmap complete
        uint32_t carry = 0;
#pragma unroll
        for (int r = 0; r < LZ_ROUNDS; r++) {
            const uint32_t idx = base + r * 32;
            uint32_t gp = __shfl_up_sync(0xffffffffu, gw[r], 1), pp = __shfl_up_sync(0xffffffffu, pw[r], 1);
            uint32_t dp = BY_KEY ? __shfl_up_sync(0xffffffffu, dw[r], 1) : 0u;
            const uint32_t gwrap = r > 0 ? __shfl_sync(0xffffffffu, gw[r > 0 ? r - 1 : 0], 31) : gb;
            const uint32_t pwrap = r > 0 ? __shfl_sync(0xffffffffu, pw[r > 0 ? r - 1 : 0], 31) : pb;
            const uint32_t dwrap = BY_KEY ? (r > 0 ? __shfl_sync(0xffffffffu, dw[r > 0 ? r - 1 : 0], 31) : db) : 0u;
            if (lane == 0) { gp = gwrap; pp = pwrap; dp = dwrap; }
            pprev[r] = pp;
            const bool hd = idx < n && idx > 0 && lz_is_head<BY_KEY>(gw[r], gp, dw[r], dp, bm, t0, idx);
            const unsigned hm = __ballot_sync(0xffffffffu, hd);
            const unsigned upto = hm & (0xffffffffu >> (31 - lane));
            const uint32_t rbase = idx - lane;
            v[r] = upto ? rbase + (31 - __clz(upto)) : carry;
            carry = hm ? rbase + (31 - __clz(hm)) : carry;
        }
        if (lane == 0) wtot[warp] = carry;
    }
    uint32_t packed[LZ_ROUNDS];  // SCATTER: digit << 16 | rank inside the warp's span
    const uint32_t sh = 8u * (lvl & 3u);
    if (SCATTER) {
#pragma unroll
        for (int r = 0; r < LZ_ROUNDS; r++) {
            const uint32_t i = base + r * 32;
            const bool valid = i < n;
            const uint32_t key = valid ? (dw[r] >> sh) & 255u : 256u + lane;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = wc[warp][key];
                wc[warp][key] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            packed[r] = (key << 16) | (old + __popc(peers & lanemask_lt()));
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- cross-tile prefixes by decoupled look-back ----
    if (SCATTER && threadIdx.x < 256) {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < LZ_WARPS; w++) {
            uint32_t t = wc[w][d];
            wc[w][d] = run;
            run += t;
        }
        st_vol(&ch.st_cnt[(size_t)tile * 256 + d], (tile == 0 ? LB_INCL : LB_AGG) | run);
        // look back eight tiles per round trip: the states are read with independent loads, then folded in order. Folding
        // aggregates instead of waiting for the predecessor's inclusive prefix keeps the tiles from forming a serial chain.
        uint32_t excl = 0;
        bool fin = false;
        for (int64_t t = (int64_t)tile - 1; t >= 0 && !fin; t -= 16) {
            uint32_t sv[16];
#pragma unroll
            for (int k = 0; k < 16; k++) sv[k] = t - k >= 0 ? ld_vol(&ch.st_cnt[(size_t)(t - k) * 256 + d]) : LB_INCL;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (fin) break;
                while ((sv[k] >> 30) == 0) sv[k] = ld_vol(&ch.st_cnt[(size_t)(t - k) * 256 + d]);
                excl += sv[k] & LB_VAL;
                if ((sv[k] >> 30) == 2u) fin = true;
            }
        }
        if (tile != 0) st_vol(&ch.st_cnt[(size_t)tile * 256 + d], LB_INCL | (excl + run));
        goff[d] = bstart[lvl * 257 + d] + excl;
    }
    if (GROUP && threadIdx.x == LZ_THREADS - 1) {
        uint32_t agg = 0;
        for (int w = 0; w < LZ_WARPS; w++) agg = max(agg, wtot[w]);
        st_vol(&ch.st_max[tile], (tile == 0 ? LB_INCL : LB_AGG) | agg);
        uint32_t pre = 0;
        bool fin = false;
        for (int64_t t = (int64_t)tile - 1; t >= 0 && !fin; t -= 8) {
            uint32_t sv[8];
#pragma unroll
            for (int k = 0; k < 8; k++) sv[k] = t - k >= 0 ? ld_vol(&ch.st_max[t - k]) : LB_INCL;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (fin) break;
                while ((sv[k] >> 30) == 0) sv[k] = ld_vol(&ch.st_max[t - k]);
                pre = max(pre, sv[k] & LB_VAL);
                if ((sv[k] >> 30) == 2u) fin = true;
            }
        }
        if (tile != 0) st_vol(&ch.st_max[tile], LB_INCL | max(pre, agg));
        s_pre = pre;
    }
    __syncthreads();
    // ---- group starts, match bookkeeping ----
    uint32_t carried[LZ_ROUNDS];
    if (GROUP) {
        uint32_t pre = s_pre;
        for (int w = 0; w < warp; w++) pre = max(pre, wtot[w]);
#pragma unroll
        for (int r = 0; r < LZ_ROUNDS; r++) {
            const uint32_t idx = base + r * 32;
            carried[r] = 0;
            if (idx < n) {
                const uint32_t g = max(pre, v[r]);
                const uint32_t p = pw[r] & LZ_POS_MASK, prev = pprev[r] & LZ_POS_MASK;
                const bool exists = g != idx && p - prev <= (uint32_t)LZ_WINDOW && lvl <= (pw[r] >> 28);
                const uint32_t word = g | (exists ? LZ_ALIVE : 0u);
                gs_final[idx] = word;
                carried[r] = word;
                if (!exists && (gw[r] & LZ_ALIVE)) match_rec[p] = (lvl - 1) << 28 | (gw[r] & LZ_GS_MASK);
                if (exists && lvl == (uint32_t)LZ_LEVELS) match_rec[p] = lvl << 28 | g;
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < LZ_ROUNDS; r++) carried[r] = gw[r];
    }
    // ---- scatter into the next level's order ----
    if (SCATTER) {
        const bool gather = (lvl & 3u) == 3u;
        uint32_t nd[LZ_ROUNDS];
        if (gather) {
#pragma unroll
            for (int r = 0; r < LZ_ROUNDS; r++) nd[r] = load4(bs + (pw[r] & LZ_POS_MASK) + lvl + 1);
        }
#pragma unroll
        for (int r = 0; r < LZ_ROUNDS; r++) {
            const uint32_t i = base + r * 32;
            if (i < n) {
                const uint32_t d = packed[r] >> 16, rk = packed[r] & 0xffffu;
                const uint32_t dst = goff[d] + wc[warp][d] + rk;
                pos_out[dst] = pw[r];
                gs_out[dst] = carried[r];
                dig_out[dst] = gather ? nd[r] : dw[r];
            }
        }
    }
}

// ---- greedy parse: orbit over bestlen[] (orbit.cuh) -------------------------
struct LzStep {
    __device__ static uint32_t step(uint32_t c) { return c >= (uint32_t)LZ_MINLEN ? c : 1u; }
    __device__ static uint32_t weight(uint32_t c) { return c >= (uint32_t)LZ_MINLEN ? 21u : 9u; }  // bits per token
};
struct LzVisit {
    const OrbitSeg* segs;
    uint32_t* bitcum;
    __device__ void operator()(uint32_t sg, uint32_t pos, uint32_t cum, uint32_t) const { bitcum[segs[sg].off + pos] = cum; }
};

// ---- token emission ----------------------------------------------------------
__global__ void __launch_bounds__(256) lz_pack_k(const uint8_t* __restrict__ bs, uint32_t n, const uint32_t* __restrict__ fs, uint32_t F,
                                                 const uint32_t* __restrict__ match_rec, const uint32_t* __restrict__ bitcum, APtrs A, uint32_t na,
                                                 const uint32_t* __restrict__ wbase, uint32_t* __restrict__ out_words) {
    const uint32_t f = blockIdx.y;  // one grid row per frame: no search for the frame of a position
    const uint32_t i = fs[f] + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fs[f + 1]) return;
    const uint32_t rel = bitcum[i];
    if (rel == EMPTY32) return;
    const uint32_t rec = match_rec[i];
    uint32_t l = rec >> 28, v, nb;
    if (l >= (uint32_t)LZ_MINLEN && (rec & LZ_RES)) {   // resolved by lz_small_k
        v = ((rec & 0xFFFFu) << 1) | (l << 17);
        nb = 21;
    } else if (l >= (uint32_t)LZ_MINLEN) {
        const uint32_t* a = A.a[l];
        const uint32_t* g = A.gs[l];
        const uint32_t glo = rec & (LZ_RES - 1u);  // start of i's level-l group in A[l]; positions ascend inside the group
        const uint32_t target = i > (uint32_t)LZ_WINDOW ? i - (uint32_t)LZ_WINDOW : 0u;
        // first j >= glo with f(j) = (j outside the group) || (a[j] >= target). f is monotone, and the answer lies inside
        // the group because i itself is a member with a[.] = i >= target. Gallop for an upper bound, then bisect.
        // (na = number of elements of the level arrays: only the large groups get that far)
        auto f = [&](uint32_t j) { return j >= na || (g[j] & LZ_GS_MASK) != glo || (a[j] & LZ_POS_MASK) >= target; };
        // Positions inside a group increase by at least 1 per index, so U = glo + (target - a[glo]) already satisfies f;
        // in a dense run (the common large group) it is the exact answer: probe U-1 first. Otherwise gallop up from
        // the group start (small or sparse groups end within a few probes), then bisect.
        const uint32_t first = a[glo] & LZ_POS_MASK;
        uint32_t lo = glo, hi = glo;
        if (first < target) {
            const uint32_t U = glo + (target - first);
            lo = glo + 1;
            hi = U;
            if (hi > lo && !f(hi - 1)) lo = hi;
            else if (hi > lo) {
                hi = hi - 1;  // f(U-1) holds
                uint32_t step = 1, probe = lo;
                while (probe < hi && !f(probe)) { lo = probe + 1; probe += step; step <<= 1; }
                if (probe < hi) hi = probe;
                while (lo < hi) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (f(mid)) hi = mid; else lo = mid + 1;
                }
            }
        }
        uint32_t off = i - (a[lo] & LZ_POS_MASK);
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

__global__ void __launch_bounds__(1024) lz_finalize_k(uint32_t F, uint32_t stub, const uint32_t* __restrict__ total_bits,
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
            uint32_t ob = total_bits[f];
            uint32_t cs = (uint32_t)((float)(int)ob / 8.0f);  // src/agmv_encode.c:176
            outbits[f] = ob;
            csize[f] = cs;
            len = 24u + stub + cs;  // AGFC hdr 16 + payload + 8 x 0xFF (+ empty AGAC chunk 8)
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
                                                         uint32_t stub, uint8_t* __restrict__ image, uint32_t audio_chunk = 0,
                                                         const uint8_t* __restrict__ audio = nullptr, uint64_t audio_size = 0) {
    const uint32_t f = blockIdx.y;
    const uint32_t cs = csize[f], len = 24u + stub + cs, usize = fs[f + 1] - fs[f];
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
            // AGMV_EncodeAudioChunk (:707-717): 'AGAC', size, then size bytes from atsample[start_point++]; frame number g of the
            // sequence starts at g * size. Reads past the track (the reference walks off its malloc) are defined as 0.
            const uint32_t k = b - 24 - cs;
            if (k < 4) v = (uint8_t)(0x43414741u /* "AGAC" */ >> (k * 8));
            else if (k < 8) v = (uint8_t)(audio_chunk >> ((k - 4) * 8));
            else {
                const uint64_t idx = (uint64_t)(first_frame_count + f) * audio_chunk + (k - 8);
                v = idx < audio_size ? audio[idx] : 0;
            }
        }
        dst[b] = v;
    }
}

// =====================================================================================================
// Large groups whose key is three equal bytes (long runs): run tables instead of sorting.
// (tools/run_path_prototype.py states the rule and checks it against the brute-force search.)
//
// For a member y of such a group let r(y) be the length of the run of that byte starting at y. A candidate x matches
// min(r(x), r(y)) bytes unless r(x) == r(y), in which case the bytes after the two runs decide. So
//   * the group's members are cut into runs (consecutive positions): start position and length per run, plus, for every
//     v = 4..15, the sorted list of the runs at least v long;
//   * y's match is min(r(y), 15, bytes left, largest r in the window), its start the earliest window member with at least
//     that r - the member of a run with r >= v that comes first is max(run start, window start), so both are answered
//     per run, by binary search, without touching the members (rg_match_k);
//   * only members with r(y) <= 14 (the last twelve of every run) can do better, through a candidate with the same r whose
//     continuation also matches: they alone stay in the compacted array that goes through the global levels, and
//     rg_match_k keeps that result when it is longer than r(y).
// A run of R members thus contributes 12 elements to the sorted passes instead of R.
// =====================================================================================================
struct RunView {
    const uint32_t* a3;
    const uint32_t* gs3;
    const uint32_t* dig3;   // key words (bytes 0..3) in level-3 order
    uint32_t* mi;           // exclusive count of members; after rg_build_k bit 31 = "is a member" (the key words are gone by then)
    const uint32_t* rsx;    // exclusive count of run starts
    const uint32_t* rs;     // per run: position of its first member
    const uint32_t* rmi;    // per run: member index of its first member (rmi[K] = number of members)
    uint32_t n;
    bool late;              // after the global levels: dig3 has been overwritten, use the flag in mi
    __device__ bool member(uint32_t i) const {
        if (late) return mi[i] >> 31;
        const uint32_t d = dig3[i];
        return ((d ^ (d >> 8)) & 0xFFFFu) == 0 && sg_group_is_large(gs3, n, gs3[i] & LZ_GS_MASK);
    }
    __device__ bool run_start(uint32_t i) const {
        if (!member(i)) return false;
        if (i == (gs3[i] & LZ_GS_MASK)) return true;
        return (a3[i] & LZ_POS_MASK) != (a3[i - 1] & LZ_POS_MASK) + 1u;
    }
    __device__ uint32_t run_of(uint32_t i) const { return rsx[i] + (run_start(i) ? 1u : 0u) - 1u; }   // i must be a member
    __device__ uint32_t run_len(uint32_t k) const { return rmi[k + 1] - rmi[k] + 2u; }               // bytes of the run
};
struct RunMemberFlag { RunView v; __device__ uint32_t operator()(uint32_t i) const { return v.member(i) ? 1u : 0u; } };
struct RunStartFlag { RunView v; __device__ uint32_t operator()(uint32_t i) const { return v.run_start(i) ? 1u : 0u; } };
struct RunLenFlag { const uint32_t* rmi; uint32_t len; __device__ uint32_t operator()(uint32_t k) const { return rmi[k + 1] - rmi[k] + 2u >= len ? 1u : 0u; } };

__global__ void __launch_bounds__(256) rg_build_k(RunView v, uint32_t* __restrict__ rs, uint32_t* __restrict__ rmi, uint32_t n_members, uint32_t n_runs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) rmi[n_runs] = n_members;
    if (i >= v.n || !v.member(i)) return;
    const uint32_t m = v.mi[i];
    if (v.run_start(i)) {
        const uint32_t k = v.rsx[i];
        rs[k] = v.a3[i] & LZ_POS_MASK;
        rmi[k] = m;
    }
    v.mi[i] = m | 0x80000000u;
}
__global__ void __launch_bounds__(256) rg_list_k(const uint32_t* __restrict__ rmi, uint32_t n_runs, uint32_t len, const uint32_t* __restrict__ prefix,
                                                 uint32_t* __restrict__ list) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_runs && rmi[k + 1] - rmi[k] + 2u >= len) list[prefix[k]] = k;
}
// what goes through the global levels: the large groups, minus the run members that the run tables settle (r >= 15)
struct LargeKeepFlag {
    RunView v;
    __device__ uint32_t operator()(uint32_t i) const {
        if (!sg_group_is_large(v.gs3, v.n, v.gs3[i] & LZ_GS_MASK)) return 0u;
        if (!v.member(i)) return 1u;
        const uint32_t k = v.run_of(i);
        const uint32_t r = v.rs[k] + v.run_len(k) - (v.a3[i] & LZ_POS_MASK);
        return r < (uint32_t)LZ_MAXLEN ? 1u : 0u;
    }
};
__global__ void __launch_bounds__(256) rg_compact_k(RunView v, const uint32_t* __restrict__ prefix, uint32_t* __restrict__ a_out,
                                                    uint32_t* __restrict__ gs_out, uint32_t* __restrict__ dig_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n || !LargeKeepFlag{v}(i)) return;
    const uint32_t gw = v.gs3[i], o = prefix[i];
    a_out[o] = v.a3[i];
    gs_out[o] = prefix[gw & LZ_GS_MASK] | (gw & LZ_ALIVE);   // the group's first kept element (the start itself may be gone)
    dig_out[o] = v.dig3[i];
}

// earliest member x of the group with lo <= x < y and r(x) >= len; k = y's run, kf = the group's first run. EMPTY32 if none.
__device__ __forceinline__ uint32_t rg_earliest(const RunView& v, const uint32_t* __restrict__ lists, const uint32_t* __restrict__ lcnt, uint32_t kcap,
                                                uint32_t lo, uint32_t y, uint32_t k, uint32_t kf, uint32_t len) {
    // last run of [kf, k] that starts at or before lo (kf if none does)
    uint32_t a = kf, b = k + 1;
    while (b - a > 1) {
        const uint32_t mid = (a + b) >> 1;
        if (v.rs[mid] <= lo) a = mid; else b = mid;
    }
    const uint32_t klo = a;
    {
        const uint32_t s = v.rs[klo], e = s + v.run_len(klo);
        const uint32_t first = max(s, lo);
        if (klo == k) return (first < y && e - first >= len) ? first : EMPTY32;
        if (first + 3u <= e && e - first >= len) return first;     // first is still a member (first <= e - 3)
    }
    // runs klo+1 .. k-1 lie inside the window entirely
    if (len <= 3u) {
        if (klo + 1 < k) return v.rs[klo + 1];
    } else {
        const uint32_t* list = lists + (size_t)(len - 4u) * kcap;
        uint32_t lo_i = 0, hi_i = lcnt[len - 4u];
        while (lo_i < hi_i) {
            const uint32_t mid = (lo_i + hi_i) >> 1;
            if (list[mid] >= klo + 1) hi_i = mid; else lo_i = mid + 1;
        }
        if (lo_i < lcnt[len - 4u] && list[lo_i] < k) return v.rs[list[lo_i]];
    }
    // y's own run: its members before y all have a larger r than y
    const uint32_t s = v.rs[k];
    return s < y ? s : EMPTY32;
}

__global__ void __launch_bounds__(256) rg_match_k(RunView v, const uint32_t* __restrict__ lists, const uint32_t* __restrict__ lcnt, uint32_t kcap,
                                                  uint32_t* __restrict__ match_rec) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n || !v.member(i)) return;
    const uint32_t gw = v.gs3[i];
    if (!(gw & LZ_ALIVE)) return;                                   // no earlier member within the window
    const uint32_t pw = v.a3[i], y = pw & LZ_POS_MASK, cap = pw >> 28;
    const uint32_t k = v.run_of(i), kf = v.rsx[gw & LZ_GS_MASK];
    const uint32_t r = v.rs[k] + v.run_len(k) - y;
    if (r < (uint32_t)LZ_MAXLEN && (match_rec[y] >> 28) > r) return;   // a same-r candidate matched past the run (global levels)
    const uint32_t lo = y > (uint32_t)LZ_WINDOW ? y - (uint32_t)LZ_WINDOW : 0u;
    for (uint32_t len = min(min(r, (uint32_t)LZ_MAXLEN), cap); len >= (uint32_t)LZ_MINLEN; len--) {
        const uint32_t x = rg_earliest(v, lists, lcnt, kcap, lo, y, k, kf, len);
        if (x != EMPTY32) {
            match_rec[y] = len << 28 | LZ_RES | (y - x);
            return;
        }
    }
    match_rec[y] = 0;   // (not reached for an alive member; clears a partial record of the global levels otherwise)
}

// Host driver of the chain path (lzchain.cuh). Same contract as lzss_encode_batch below.
inline void lzss_encode_batch_chain(LzWork& wk, const uint8_t* bs, const uint32_t* fs, uint32_t F, uint32_t n, uint32_t ntile, uint32_t max_usize,
                                    uint32_t first_frame_count, uint8_t* image, LaunchCtx& lc) {
    cudaStream_t st = lc.st;
    KL(lc, KC_LZ_INIT, (lzc_wbase_k<<<cdiv(F + 1, 256), 256, 0, st>>>(fs, F, wk.wbase)));
    const uint32_t* lw15 = wk.lw[0];
    if (n > 0) {
        cudaMemsetAsync(wk.bitcum, 0xFF, (size_t)n * 4, st);
        // persistent walking kernels: enough warps to fill the GPU, chunks handed out through one counter per launch
        const uint32_t nb3 = std::min<uint32_t>(cdiv(n, LZC_BCHUNK), wk.link3_blocks), nb = std::min<uint32_t>(cdiv(n, LZC_BCHUNK), wk.level_blocks);
        cudaMemsetAsync(wk.counters, 0, 16 * sizeof(uint32_t), st);
        KL(lc, KC_LZ_LINK, (lzc_hashlink_k<<<wk.n_items, 32, LZC_TAB_BYTES, st>>>(bs, n, fs, wk.items, wk.lw[1], wk.rsd)));
        KL(lc, KC_LZ_LINK3, (lzc_link3_k<<<nb3, LZC_THREADS, 0, st>>>(bs, fs, F, n, wk.lw[1], wk.rsd, wk.lw[0], wk.bestlen, wk.counters)));
        int cur = 0;
        for (uint32_t L = LZ_MINLEN; L < (uint32_t)LZ_MAXLEN; L++, cur ^= 1)
            KL(lc, KC_LZ_LEVEL, (lzc_level_k<<<nb, LZC_THREADS, 0, st>>>(bs, n, L, wk.lw[cur], wk.rsd, wk.lw[cur ^ 1], wk.match_rec, wk.bestlen,
                                                                         wk.counters + (L - LZ_MINLEN + 1))));
        lw15 = wk.lw[cur];
    }
    orbit_run<LZ_MAXLEN, LzStep>(wk.bestlen, wk.segs, F, wk.seg_len, ntile, wk.orb, LzVisit{wk.segs, wk.bitcum}, lc, KC_LZ_PARSE);
    if (n > 0) {
        size_t words = (((size_t)n * 9) >> 5) + 3 * (size_t)F + 4;
        cudaMemsetAsync(wk.out_words, 0, words * 4, st);
        dim3 pgrid(cdiv(max_usize, (uint32_t)LZC_BCHUNK), F);
        KL(lc, KC_LZ_PACK, (lzc_pack_k<<<pgrid, LZC_THREADS, 0, st>>>(bs, fs, wk.bestlen, wk.match_rec, wk.bitcum, lw15, wk.rsd, wk.wbase, wk.out_words)));
    }
    KL(lc, KC_LZ_CHUNK, (lz_finalize_k<<<1, 1024, 0, st>>>(F, wk.stub_bytes, wk.orb.final_cum, wk.outbits, wk.csize, wk.chunk_off)));
    dim3 grid(32, F);
    KL(lc, KC_LZ_CHUNK, (lz_write_chunks_k<<<grid, 256, 0, st>>>(fs, wk.csize, wk.chunk_off, wk.wbase, wk.out_words, first_frame_count, wk.stub_bytes, image,
                                                                 wk.audio_chunk, wk.audio, wk.audio_size)));
}

// Host driver. bs: batch bitstream (n bytes + >=16 bytes of readable padding);
// fs: device array of F+1 frame starts (fs[0]=0, fs[F]=n); wk.segs / wk.seg_len
// describe the same frames for the parse (filled by the caller, ntile tiles in
// total). Results: wk.csize, wk.outbits, wk.chunk_off on the device and the chunk
// image written to `image` (capacity >= 32*F + 9n/8 + 8).
inline void lzss_encode_batch(LzWork& wk, const uint8_t* bs, const uint32_t* fs, uint32_t F, uint32_t n, uint32_t ntile, uint32_t max_usize,
                              uint32_t first_frame_count, uint8_t* image, LaunchCtx& lc) {
    if (!wk.legacy) { lzss_encode_batch_chain(wk, bs, fs, F, n, ntile, max_usize, first_frame_count, image, lc); return; }
    cudaStream_t st = lc.st;
    const uint32_t nthreads = 256;
    uint32_t n_large = n;                       // elements of the level arrays A[3..15] (large groups only; everything in the fused variant)
    const uint32_t* level3_a = wk.A[LZ_MINLEN];
    const uint32_t* level3_gs = wk.GS[LZ_MINLEN];
    uint32_t cover = (n + 1 > F + 1 ? n + 1 : F + 1);
    KL(lc, KC_LZ_INIT, (lz_init_k<<<cdiv(cover, nthreads), nthreads, 0, st>>>(bs, n, fs, F, wk.A[0], wk.GS[0], wk.dig4[0], wk.match_rec, wk.bitcum, wk.wbase)));
    if (n > 0) {
        const uint32_t nt = cdiv(n, RX_TILE);
        if (wk.fused) {
            // key-byte bucket starts of all 15 levels from one byte histogram
            cudaMemsetAsync(wk.bytehist, 0, 256 * 4, st);
            KL(lc, KC_RX_HIST, (lz_bytehist_k<<<std::min<uint32_t>(cdiv((size_t)n + LZ_LEVELS, 1024), 1184), 256, 0, st>>>(bs, n + LZ_LEVELS, wk.bytehist)));
            KL(lc, KC_RX_HIST, (lz_bstart_k<<<LZ_LEVELS, 256, 0, st>>>(bs, n, wk.bytehist, wk.bstart)));
            // gs_carry[0] <- frame starts (the level-0 "groups")
            cudaMemcpyAsync(wk.gs_carry[0], wk.GS[0], (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
            LzChain chain{wk.chain_mem, wk.chain_mem + 1, wk.chain_mem + 1 + nt};
            for (uint32_t lvl = 0; lvl <= (uint32_t)LZ_LEVELS; lvl++) {
                cudaMemsetAsync(wk.chain_mem, 0, ((size_t)nt * 257 + 1) * 4, st);  // ticket, st_max[nt], st_cnt[nt][256]
                uint32_t* gin = wk.gs_carry[lvl & 1];
                uint32_t* gout = wk.gs_carry[(lvl & 1) ^ 1];
                uint32_t* din = wk.dig4[lvl & 1];
                uint32_t* dout = wk.dig4[(lvl & 1) ^ 1];
                uint32_t* pout = lvl < (uint32_t)LZ_LEVELS ? wk.A[lvl + 1] : nullptr;
                if (lvl < (uint32_t)LZ_MINLEN)
                    KL(lc, KC_RX_SCATTER, (lz_level_k<false, false, true><<<nt, LZ_THREADS, 0, st>>>(bs, n, lvl, wk.A[lvl], gin, din, wk.GS[lvl], wk.match_rec, pout,
                                                                                                     gout, dout, wk.bstart, chain)));
                else if (lvl == (uint32_t)LZ_MINLEN)
                    KL(lc, KC_LZ_GROUP, (lz_level_k<true, true, true><<<nt, LZ_THREADS, 0, st>>>(bs, n, lvl, wk.A[lvl], gin, din, wk.GS[lvl], wk.match_rec, pout,
                                                                                                 gout, dout, wk.bstart, chain)));
                else if (lvl < (uint32_t)LZ_LEVELS)
                    KL(lc, KC_LZ_GROUP, (lz_level_k<true, false, true><<<nt, LZ_THREADS, 0, st>>>(bs, n, lvl, wk.A[lvl], gin, din, wk.GS[lvl], wk.match_rec, pout,
                                                                                                  gout, dout, wk.bstart, chain)));
                else
                    KL(lc, KC_LZ_GROUP, (lz_level_k<true, false, false><<<nt, LZ_THREADS, 0, st>>>(bs, n, lvl, wk.A[lvl], gin, din, wk.GS[lvl], wk.match_rec, pout,
                                                                                                   gout, dout, wk.bstart, chain)));
            }
        } else {
            // ---- levels 1..3 over every position (levels 1 and 2 never carry a match: no group phase, the frame start rides
            //      along as the group; level 3 derives its groups from (frame, 3-byte key)) ----
            KL(lc, KC_RX_HIST, (radix_hist_k<LzDigit><<<nt, RX_THREADS, 0, st>>>(LzDigit{wk.dig4[0], 0u}, n, nt, wk.tile_hist[0])));
            for (uint32_t L = 0; L < (uint32_t)LZ_MINLEN; L++) {
                uint32_t* din = wk.dig4[L & 1];
                uint32_t* dout = wk.dig4[(L & 1) ^ 1];
                uint32_t* th = wk.tile_hist[L & 1];
                uint32_t* th_next = wk.tile_hist[(L & 1) ^ 1];
                const uint32_t Lnew = L + 1;
                uint32_t* gs_dst = Lnew < (uint32_t)LZ_MINLEN ? wk.GS[Lnew] : wk.gs_tmp;
                device_scan<SumOp, true>(LoadU32{th}, StoreU32{th}, 256u * nt, wk.scan_ws, lc, KC_RX_SCAN);
                KL(lc, KC_RX_SCATTER, (lz_scatter_k<<<nt, LZ_THREADS, LZ_SCATTER_SMEM, st>>>(bs, wk.A[L], wk.GS[L], din, wk.A[Lnew], gs_dst, dout, L, n, nt, th)));
                if (Lnew < (uint32_t)LZ_MINLEN) {
                    KL(lc, KC_RX_HIST, (radix_hist_k<LzDigit><<<nt, RX_THREADS, 0, st>>>(LzDigit{dout, 8u * (Lnew & 3u)}, n, nt, th_next)));
                    continue;
                }
                KL(lc, KC_LZ_GROUP, (lz_group_reduce_k<true><<<nt, LZ_THREADS, 0, st>>>(wk.gs_tmp, th, nt, n, dout, wk.scan_ws)));
                KL(lc, KC_LZ_GROUP, (scan_partials_k<MaxOp><<<1, 1024, 0, st>>>(wk.scan_ws, nt)));
                KL(lc, KC_LZ_GROUP, (lz_group_apply_k<true, false><<<nt, LZ_THREADS, 0, st>>>(wk.gs_tmp, th, nt, n, wk.scan_ws, wk.A[Lnew], wk.GS[Lnew],
                                                                                                wk.match_rec, Lnew, dout, th_next)));
            }
            // ---- tiny groups: all pairs; small groups: levels 4..15 in shared memory; both resolve to (length, offset) ----
            KL(lc, KC_LZ_TINY, (lz_tiny_k<<<cdiv(n, 256u), 256, 0, st>>>(bs, n, wk.A[LZ_MINLEN], wk.GS[LZ_MINLEN], wk.match_rec)));
            KL(lc, KC_LZ_SMALL, (lz_small_k<<<cdiv(n, (uint32_t)SG_W), SG_THREADS, SG_SMEM, st>>>(bs, n, wk.A[LZ_MINLEN], wk.GS[LZ_MINLEN], wk.match_rec)));
            // ---- large groups: compact them (A[1] / GS[1] are free again) and run the global levels on what is left ----
            uint32_t* dig3 = wk.dig4[LZ_MINLEN & 1];          // key words in level-3 order
            uint32_t* digc = wk.dig4[(LZ_MINLEN & 1) ^ 1];
            uint32_t* prefix = wk.gs_tmp;
            // run tables of the large three-equal-byte groups (A[2] / GS[2] are free: member and run-start counts)
            RunView rv{wk.A[LZ_MINLEN], wk.GS[LZ_MINLEN], dig3, wk.A[2], wk.GS[2], nullptr, nullptr, n, false};
            uint32_t n_runs = 0, n_members = 0, kcap = 0;
            uint32_t *lists = nullptr, *lcnt = nullptr;
            bool use_runs = false;
            if (wk.runs) {
                const uint32_t st_tiles = cdiv(n, (uint32_t)SCAN_TILE);
                device_scan<SumOp, true>(RunMemberFlag{rv}, StoreU32{wk.A[2]}, n, wk.scan_ws, lc, KC_LZ_GROUP);
                cudaMemcpyAsync(&n_members, wk.scan_ws + st_tiles, 4, cudaMemcpyDeviceToHost, st);
                device_scan<SumOp, true>(RunStartFlag{rv}, StoreU32{wk.GS[2]}, n, wk.scan_ws + st_tiles + 2, lc, KC_LZ_GROUP);
                cudaMemcpyAsync(&n_runs, wk.scan_ws + st_tiles + 2 + st_tiles, 4, cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
                // worth it only for long runs, and the tables have to fit: rs, rmi, 12 lists, 16 counters
                kcap = n_runs + 8;
                use_runs = n_runs > 0 && (uint64_t)n_runs * 16u <= n_members && (uint64_t)kcap * 14u + 32u <= wk.cap_n;
            }
            if (use_runs) {
                uint32_t* rs = wk.run_ws;
                uint32_t* rmi = rs + kcap;
                lcnt = rmi + kcap;
                lists = lcnt + 16;
                rv.rs = rs;
                rv.rmi = rmi;
                KL(lc, KC_LZ_GROUP, (rg_build_k<<<cdiv(n, nthreads), nthreads, 0, st>>>(rv, rs, rmi, n_members, n_runs)));
                for (uint32_t len = 4; len <= (uint32_t)LZ_MAXLEN; len++) {
                    uint32_t* pre = wk.tile_hist[0];   // n_runs <= n / 16 words: fits the tile histogram buffer
                    uint32_t* ws = wk.scan_ws;
                    device_scan<SumOp, true>(RunLenFlag{rmi, len}, StoreU32{pre}, n_runs, ws, lc, KC_LZ_GROUP);
                    cudaMemcpyAsync(lcnt + (len - 4), ws + cdiv(n_runs, (uint32_t)SCAN_TILE), 4, cudaMemcpyDeviceToDevice, st);
                    KL(lc, KC_LZ_GROUP, (rg_list_k<<<cdiv(n_runs, nthreads), nthreads, 0, st>>>(rmi, n_runs, len, pre, lists + (size_t)(len - 4) * kcap)));
                }
                device_scan<SumOp, true>(LargeKeepFlag{rv}, StoreU32{prefix}, n, wk.scan_ws, lc, KC_LZ_GROUP);
            } else {
                device_scan<SumOp, true>(LargeFlag{wk.GS[LZ_MINLEN], n}, StoreU32{prefix}, n, wk.scan_ws, lc, KC_LZ_GROUP);
            }
            cudaMemcpyAsync(&n_large, wk.scan_ws + cdiv(n, (uint32_t)SCAN_TILE), 4, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            if (n_large > 0) {
                const uint32_t m = n_large, mt = cdiv(m, RX_TILE);
                if (use_runs) KL(lc, KC_LZ_GROUP, (rg_compact_k<<<cdiv(n, nthreads), nthreads, 0, st>>>(rv, prefix, wk.A[1], wk.GS[1], digc)));
                else KL(lc, KC_LZ_GROUP, (lz_compact_large_k<<<cdiv(n, nthreads), nthreads, 0, st>>>(wk.A[LZ_MINLEN], wk.GS[LZ_MINLEN], dig3, prefix, n, wk.A[1], wk.GS[1], digc)));
                level3_a = wk.A[1];
                level3_gs = wk.GS[1];
                // dig parity is flipped from here on: level L reads dig4[(L & 1) ^ 1]
                KL(lc, KC_RX_HIST, (radix_hist_k<LzDigit><<<mt, RX_THREADS, 0, st>>>(LzDigit{digc, 8u * (LZ_MINLEN & 3u)}, m, mt, wk.tile_hist[LZ_MINLEN & 1])));
                for (uint32_t L = LZ_MINLEN; L < (uint32_t)LZ_LEVELS; L++) {
                    uint32_t* din = wk.dig4[(L & 1) ^ 1];
                    uint32_t* dout = wk.dig4[L & 1];
                    uint32_t* th = wk.tile_hist[L & 1];          // counts of this level's key byte
                    uint32_t* th_next = wk.tile_hist[(L & 1) ^ 1];
                    const uint32_t Lnew = L + 1;
                    const uint32_t* a_in = L == (uint32_t)LZ_MINLEN ? level3_a : wk.A[L];
                    const uint32_t* g_in = L == (uint32_t)LZ_MINLEN ? level3_gs : wk.GS[L];
                    device_scan<SumOp, true>(LoadU32{th}, StoreU32{th}, 256u * mt, wk.scan_ws, lc, KC_RX_SCAN);
                    KL(lc, KC_RX_SCATTER, (lz_scatter_k<<<mt, LZ_THREADS, LZ_SCATTER_SMEM, st>>>(bs, a_in, g_in, din, wk.A[Lnew], wk.gs_tmp, dout, L, m, mt, th)));
                    KL(lc, KC_LZ_GROUP, (lz_group_reduce_k<false><<<mt, LZ_THREADS, 0, st>>>(wk.gs_tmp, th, mt, m, dout, wk.scan_ws)));
                    KL(lc, KC_LZ_GROUP, (scan_partials_k<MaxOp><<<1, 1024, 0, st>>>(wk.scan_ws, mt)));
                    if (Lnew < (uint32_t)LZ_LEVELS)
                        KL(lc, KC_LZ_GROUP, (lz_group_apply_k<false, true><<<mt, LZ_THREADS, 0, st>>>(wk.gs_tmp, th, mt, m, wk.scan_ws, wk.A[Lnew], wk.GS[Lnew],
                                                                                                        wk.match_rec, Lnew, dout, th_next)));
                    else
                        KL(lc, KC_LZ_GROUP, (lz_group_apply_k<false, false><<<mt, LZ_THREADS, 0, st>>>(wk.gs_tmp, th, mt, m, wk.scan_ws, wk.A[Lnew], wk.GS[Lnew],
                                                                                                         wk.match_rec, Lnew, dout, th_next)));
                }
            }
            if (use_runs) {
                rv.late = true;
                KL(lc, KC_LZ_GROUP, (rg_match_k<<<cdiv(n, nthreads), nthreads, 0, st>>>(rv, lists, lcnt, kcap, wk.match_rec)));
            }
        }
        KL(lc, KC_LZ_GROUP, (lz_bestlen_k<<<cdiv(n, nthreads), nthreads, 0, st>>>(wk.match_rec, n, wk.bestlen)));
    }
    orbit_run<LZ_MAXLEN, LzStep>(wk.bestlen, wk.segs, F, wk.seg_len, ntile, wk.orb, LzVisit{wk.segs, wk.bitcum}, lc, KC_LZ_PARSE);
    if (n > 0) {
        size_t words = (((size_t)n * 9) >> 5) + 3 * (size_t)F + 4;
        cudaMemsetAsync(wk.out_words, 0, words * 4, st);
        APtrs ap;
        for (int l = 0; l <= LZ_LEVELS; l++) { ap.a[l] = wk.A[l]; ap.gs[l] = wk.GS[l]; }
        ap.a[LZ_MINLEN] = level3_a;
        ap.gs[LZ_MINLEN] = level3_gs;
        dim3 pgrid(cdiv(max_usize, nthreads), F);
        KL(lc, KC_LZ_PACK, (lz_pack_k<<<pgrid, nthreads, 0, st>>>(bs, n, fs, F, wk.match_rec, wk.bitcum, ap, n_large, wk.wbase, wk.out_words)));
    }
    KL(lc, KC_LZ_CHUNK, (lz_finalize_k<<<1, 1024, 0, st>>>(F, wk.stub_bytes, wk.orb.final_cum, wk.outbits, wk.csize, wk.chunk_off)));
    dim3 grid(32, F);
    KL(lc, KC_LZ_CHUNK, (lz_write_chunks_k<<<grid, 256, 0, st>>>(fs, wk.csize, wk.chunk_off, wk.wbase, wk.out_words, first_frame_count, wk.stub_bytes, image,
                                                                 wk.audio_chunk, wk.audio, wk.audio_size)));
}

}  // namespace agmvb
