// Tile-parallel resolution of a forward pointer chase.
//
// Several stages of this codec are "start at 0, repeatedly jump ahead by a
// data-dependent step of at most S bytes": the greedy LZSS parse (step = match
// length or 1), and the decoder's walk over variable-length block records
// (step = record length). Which positions get visited is inherently serial,
// but a tile of T positions can only be entered at one of its first S
// offsets, so:
//   1. orbit_spec_k  finds, for each of the S entry offsets of a tile, where a
//                    walk from there leaves the tile and the weight it gathers.
//                    One lane owns one tile and sweeps it BACKWARDS once: the
//                    answer for position p is the answer for p + step(p) (or p's
//                    own, if that already lies past the tile), kept in a ring of
//                    the last few positions - one step per position instead of
//                    one walk per entry offset;
//   2. orbit_chain_k chains the tiles of each segment (a few hundred table
//                    look-ups, one warp per segment, tables staged in smem);
//   3. orbit_mark_k  re-walks each tile from its true entry (again a lane per
//                    tile) and hands every visited position and its running
//                    weight to a visitor.
// Segments are independent walks (one per frame). `code[p]` is one byte per
// position; Dec::step / Dec::weight decode it.
#pragma once
#include "common.cuh"

namespace agmvb {

constexpr int ORB_TILE = 1024;
constexpr int ORB_SP = 40;  // padded number of entry offsets per tile (S <= 40)

struct OrbitSeg {
    uint64_t off;        // offset of the segment's first position in the code array
    uint32_t cap_len;    // upper bound of the segment length (tile count is derived from it on the host)
    uint32_t tile_base;  // index of the segment's first tile in the batch-wide tile tables
};

struct OrbitTables {
    uint8_t* exit_tab = nullptr;   // [tile][ORB_SP]
    uint16_t* w_tab = nullptr;     // [tile][ORB_SP]
    uint8_t* entry_tab = nullptr;  // [tile]
    uint32_t* cumbase = nullptr;   // [tile]
    uint32_t* final_pos = nullptr; // [segment] position (relative to the segment) where the walk left the segment
    uint32_t* final_cum = nullptr; // [segment] total weight
};

__device__ __forceinline__ uint32_t orbit_seg_of(const OrbitSeg* __restrict__ segs, uint32_t nseg, uint32_t tile) {
    uint32_t lo = 0, hi = nseg;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (segs[mid].tile_base <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// byte `i` of a stream read through aligned 32-bit words (w = the word that holds it, reloaded when i crosses into another word)
struct OrbitBytes {
    const uint32_t* words;   // aligned base
    uint32_t shift0;         // byte offset of stream position 0 inside words[0]
    uint32_t w, widx;
    __device__ __forceinline__ void init(const uint8_t* base) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(base);
        words = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        shift0 = (uint32_t)(a & 3u);
        widx = 0xFFFFFFFFu;
        w = 0;
    }
    __device__ __forceinline__ uint32_t at(uint32_t i) {
        const uint32_t b = i + shift0;
        if ((b >> 2) != widx) { widx = b >> 2; w = __ldg(words + widx); }
        return (w >> ((b & 3u) * 8u)) & 0xFFu;
    }
};

constexpr int ORB_LANES = 128;   // tiles per block (one lane each)

template <int S, class Dec>
__global__ void __launch_bounds__(ORB_LANES) orbit_spec_k(const uint8_t* __restrict__ code, const OrbitSeg* __restrict__ segs, uint32_t nseg,
                                                          const uint32_t* __restrict__ seg_len, uint32_t ntile, OrbitTables tb) {
    constexpr int RING = S <= 15 ? 16 : 64;              // >= S + 1 and a power of two
    __shared__ uint32_t ring[RING][ORB_LANES];           // [p & (RING - 1)][lane]: exit << 24 | weight of a walk that starts at p
    const uint32_t tile = blockIdx.x * ORB_LANES + threadIdx.x;
    if (tile >= ntile) return;
    const uint32_t sg = orbit_seg_of(segs, nseg, tile);
    const OrbitSeg seg = segs[sg];
    const uint32_t len = seg_len[sg];
    const uint32_t t0 = (tile - seg.tile_base) * ORB_TILE;
    if (t0 >= len) return;  // tile beyond the actual segment length: never reached by the chain
    const uint32_t end = min(t0 + ORB_TILE, len);
    OrbitBytes c;
    c.init(code + seg.off);
    for (uint32_t i = end; i-- > t0;) {
        const uint32_t cc = c.at(i);
        const uint32_t nxt = i + Dec::step(cc);
        uint32_t v = nxt >= end ? (nxt - end) << 24 : ring[nxt & (RING - 1)][threadIdx.x];
        v += Dec::weight(cc);
        ring[i & (RING - 1)][threadIdx.x] = v;
    }
    for (int e = 0; e < S; e++) {
        const uint32_t i = t0 + e;
        // an entry offset at or past the end of a short last tile leaves it at once
        const uint32_t v = i < end ? ring[i & (RING - 1)][threadIdx.x] : (i - end) << 24;
        tb.exit_tab[(size_t)tile * ORB_SP + e] = (uint8_t)(v >> 24);
        tb.w_tab[(size_t)tile * ORB_SP + e] = (uint16_t)(v & 0xFFFFu);
    }
}

// one warp per segment
template <int S>
__global__ void __launch_bounds__(128) orbit_chain_k(const OrbitSeg* __restrict__ segs, uint32_t nseg, const uint32_t* __restrict__ seg_len,
                                                     OrbitTables tb) {
    constexpr int CH = 32;  // tiles staged per round
    __shared__ uint8_t se[4][CH * ORB_SP];
    __shared__ uint16_t sw[4][CH * ORB_SP];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t sg = blockIdx.x * 4 + warp;
    if (sg >= nseg) return;
    const OrbitSeg seg = segs[sg];
    const uint32_t len = seg_len[sg];
    const uint32_t nt = (len + ORB_TILE - 1) / ORB_TILE;
    uint32_t e = 0, cum = 0;
    for (uint32_t c0 = 0; c0 < nt; c0 += CH) {
        const uint32_t cnt = min((uint32_t)CH, nt - c0);
        const size_t base = (size_t)(seg.tile_base + c0) * ORB_SP;
        __syncwarp();
        for (uint32_t k = lane; k < cnt * ORB_SP; k += 32) { se[warp][k] = tb.exit_tab[base + k]; sw[warp][k] = tb.w_tab[base + k]; }
        __syncwarp();
        if (lane == 0) {
            for (uint32_t k = 0; k < cnt; k++) {
                tb.entry_tab[seg.tile_base + c0 + k] = (uint8_t)e;
                tb.cumbase[seg.tile_base + c0 + k] = cum;
                const uint32_t t0 = (c0 + k) * ORB_TILE;
                if (t0 + e >= len) { /* entered past the end: nothing to walk */ }
                cum += sw[warp][k * ORB_SP + e];
                e = se[warp][k * ORB_SP + e];
            }
        }
    }
    if (lane == 0) {
        tb.final_pos[sg] = len + e;  // nt == 0 (empty segment): position 0
        tb.final_cum[sg] = cum;
    }
}

// Visitor: void operator()(uint32_t seg, uint32_t pos_in_seg, uint32_t cum_before, uint32_t code)
template <int S, class Dec, class Visit>
__global__ void __launch_bounds__(ORB_LANES) orbit_mark_k(const uint8_t* __restrict__ code, const OrbitSeg* __restrict__ segs, uint32_t nseg,
                                                          const uint32_t* __restrict__ seg_len, uint32_t ntile, OrbitTables tb, Visit visit) {
    const uint32_t tile = blockIdx.x * ORB_LANES + threadIdx.x;
    if (tile >= ntile) return;
    const uint32_t sg = orbit_seg_of(segs, nseg, tile);
    const OrbitSeg seg = segs[sg];
    const uint32_t len = seg_len[sg];
    const uint32_t t0 = (tile - seg.tile_base) * ORB_TILE;
    if (t0 >= len) return;
    const uint32_t end = min(t0 + ORB_TILE, len);
    OrbitBytes c;
    c.init(code + seg.off);
    uint32_t i = t0 + tb.entry_tab[tile], cum = tb.cumbase[tile];
    while (i < end) {
        const uint32_t cc = c.at(i);
        visit(sg, i, cum, cc);
        cum += Dec::weight(cc);
        i += Dec::step(cc);
    }
}

// host-side: number of tiles of a segment of (at most) `cap_len` positions
inline uint32_t orbit_tiles(uint32_t cap_len) { return (cap_len + ORB_TILE - 1) / ORB_TILE; }

// steps 1 and 2 only: entry offset and running weight of every tile (tb.entry_tab / tb.cumbase), end of every segment
template <int S, class Dec>
inline void orbit_prepare(const uint8_t* code, const OrbitSeg* d_segs, uint32_t nseg, const uint32_t* d_seg_len, uint32_t ntile, OrbitTables tb,
                          LaunchCtx& lc, int cls) {
    static_assert(S <= ORB_SP, "too many entry offsets");
    if (nseg == 0) return;
    if (ntile) KL(lc, cls, (orbit_spec_k<S, Dec><<<cdiv(ntile, ORB_LANES), ORB_LANES, 0, lc.st>>>(code, d_segs, nseg, d_seg_len, ntile, tb)));
    KL(lc, cls, (orbit_chain_k<S><<<cdiv(nseg, 4), 128, 0, lc.st>>>(d_segs, nseg, d_seg_len, tb)));
}

template <int S, class Dec, class Visit>
inline void orbit_run(const uint8_t* code, const OrbitSeg* d_segs, uint32_t nseg, const uint32_t* d_seg_len, uint32_t ntile, OrbitTables tb,
                      Visit visit, LaunchCtx& lc, int cls) {
    if (nseg == 0) return;
    orbit_prepare<S, Dec>(code, d_segs, nseg, d_seg_len, ntile, tb, lc, cls);
    if (ntile) KL(lc, cls, (orbit_mark_k<S, Dec, Visit><<<cdiv(ntile, ORB_LANES), ORB_LANES, 0, lc.st>>>(code, d_segs, nseg, d_seg_len, ntile, tb, visit)));
}

}  // namespace agmvb
