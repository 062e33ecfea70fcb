// K4 match finder, per-position logic (reference: AGMV_LZSS, src/agmv_encode.c:101-177).
//
// The reference finds, at every parse position, the longest match (3..15 bytes, overlap allowed) inside the previous
// 65535 bytes and, among equal lengths, the EARLIEST start, by brute force. Here the same answer comes from
// "occurrence chains", one level per match length:
//
//   dist_L[p] = distance from p back to the most recent earlier position of the same frame whose first L bytes equal
//               p's, if that distance is <= 65535; 0 if there is none.
//
//   * a match of length >= L exists for p  <=>  dist_L[p] != 0, so bestlen[p] = the last L for which that holds;
//   * p, p - dist_L[p], (p - dist_L[p]) - dist_L[p - dist_L[p]], ... enumerates, most recent first, every earlier
//     occurrence of p's L-gram inside the window (a link that is 0 is farther than 65535 from its own position, hence
//     from every later one);
//   * dist_{L+1}[p] is found by walking p's level-L chain until an element whose byte L equals p's byte L: the first
//     such element is the most recent occurrence of the (L+1)-gram. Typically one or two hops;
//   * a position whose walk runs off the window has no longer match; the walk has then visited its whole level-L chain,
//     and the LAST element it saw is the earliest occurrence inside the window - the start the reference picks.
//
// So no position is ever sorted or moved: a level reads one 16-bit link per position and writes one. The level-3 links
// come from a hashed "previous occurrence" table (lzchain.cuh: lzc_hashlink_k) refined by lzc_link3_hop below.
//
// Byte runs are the one input on which a chain walk degenerates (every position of a long run is on the chain of every
// other one). They are skipped exactly: when the walk stands on an element k whose link is 1, k-1 .. runstart(k) all carry
// the same L-gram (all one byte b) and the same byte L (b again), so if b is not the byte looked for none of them can
// end the walk, and the walk continues from the run's first position.
//
// A walk is taken one hop at a time against ONE TILE of the frame held in fast memory (the links and the bytes of
// positions [t0, t0 + tile)): a hop whose target lies below the tile returns LZC_LEAVE with the walker unchanged, to be
// resumed when that lower tile is resident (lzchain.cuh processes a frame's tiles from the last to the first).
//
// This header is plain C++ (no CUDA types): the kernels in lzchain.cuh drive these functions, and
// tests/lzchain_host_check.cpp compiles them for the host to check the logic against a brute-force search.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define LZC_HD __host__ __device__ __forceinline__
#else
#define LZC_HD static inline
#endif

namespace agmvb {

constexpr uint32_t LZC_WINDOW = 65535u;     // src/agmv_encode.c:102
constexpr uint32_t LZC_RESOLVED = 1u << 27; // match_rec: length << 28 | LZC_RESOLVED | offset

LZC_HD uint32_t lzc_hash(uint32_t gram24, int bits) { return (gram24 * 2654435761u) >> (32 - bits); }

enum { LZC_GO = 0, LZC_FOUND = 1, LZC_END = 2, LZC_LEAVE = 3 };

// A walk in progress. p = position (frame-relative), acc = distance already covered (p - acc is the chain element the walk
// stands on; acc == 0: still on p itself), dist = link of that element (the next hop), key:
//   level walk: byte L of p | ext << 8 | neq << 9   (ext: a match of length L+1 fits before the frame end - the reference
//               caps matches there, src/agmv_encode.c:121-123; neq: byte L of p differs from byte L-1 of p)
//   link3 walk: the three bytes at p
struct LzcWalker { uint32_t p, acc, dist, key; };

LZC_HD uint32_t lzc_level_key(uint32_t byte_l, uint32_t byte_before, bool ext) { return byte_l | (ext ? 1u << 8 : 0u) | (byte_l != byte_before ? 1u << 9 : 0u); }

// One hop of a level-L walk. sd / sraw: links and bytes of the resident tile, indexed by (position - t0); sraw reaches 32
// bytes past the tile. g_dl / g_rsd: the frame's whole link / run-start-distance arrays (used by the run skip only).
//   LZC_FOUND: w.acc = dist_{L+1}[p].   LZC_END: no match of length L+1; w.acc = offset of the EARLIEST level-L occurrence
//   inside the window, i.e. of p's final match of length L.   LZC_LEAVE: resume in the tile that holds p - w.acc - w.dist.
LZC_HD int lzc_level_hop(LzcWalker& w, uint32_t L, uint32_t t0, const uint16_t* sd, const uint8_t* sraw, const uint16_t* g_dl, const uint16_t* g_rsd) {
    const uint32_t nacc = w.acc + w.dist;
    if (nacc > LZC_WINDOW) return LZC_END;
    const uint32_t k = w.p - nacc;
    if (k < t0) return LZC_LEAVE;
    w.acc = nacc;
    const bool ext = (w.key >> 8) & 1u, neq = (w.key >> 9) & 1u;
    if (ext && sraw[k - t0 + L] == (w.key & 0xFFu)) return LZC_FOUND;
    w.dist = sd[k - t0];
    if (w.dist == 1u && (!ext || neq)) {
        // k stands inside a run of one byte b (its L-gram, hence p's, is all b, and neq says p's byte L is not b):
        // everything back to the run start is on the chain and carries byte L == b
        const uint32_t r = g_rsd[k];
        if (w.acc + r > LZC_WINDOW) { w.acc = LZC_WINDOW; return LZC_END; }   // the window ends inside the run: p - 65535 is on the chain
        w.acc += r;
        const uint32_t s = k - r;
        w.dist = s >= t0 ? sd[s - t0] : g_dl[s];
    }
    return w.dist ? LZC_GO : LZC_END;
}

// One hop of a hash-chain walk towards the most recent earlier position with p's three bytes (level 3).
//   LZC_FOUND: w.acc = dist_3[p].   LZC_END: none inside the window.
LZC_HD int lzc_link3_hop(LzcWalker& w, uint32_t t0, const uint16_t* sd, const uint8_t* sraw, const uint16_t* g_dh, const uint16_t* g_rsd) {
    const uint32_t nacc = w.acc + w.dist;
    if (nacc > LZC_WINDOW) return LZC_END;
    const uint32_t k = w.p - nacc;
    if (k < t0) return LZC_LEAVE;
    w.acc = nacc;
    const uint8_t* q = sraw + (k - t0);
    const uint32_t b0 = q[0], b1 = q[1], b2 = q[2];
    if ((b0 | b1 << 8 | b2 << 16) == w.key) return LZC_FOUND;
    w.dist = sd[k - t0];
    // run skip: k reads b,b,b and so does k-1 (its link is 1): every position back to the run start has the same gram,
    // hence the same hash, and is the next chain element; none of them is p's gram (k was not)
    if (w.dist == 1u && b0 == b1 && b1 == b2) {
        const uint32_t r = g_rsd[k];
        if (r) {
            if (w.acc + r > LZC_WINDOW) return LZC_END;
            w.acc += r;
            const uint32_t s = k - r;
            w.dist = s >= t0 ? sd[s - t0] : g_dh[s];
        }
    }
    return w.dist ? LZC_GO : LZC_END;
}

// positions whose match reaches 15 bytes: one hop towards the end of the level-15 chain; on LZC_END w.acc is the offset
// of the earliest occurrence of the 15-gram inside the window. (g15 / g_rsd: whole-frame arrays, w.key unused.)
LZC_HD int lzc_end_hop(LzcWalker& w, const uint16_t* g15, const uint16_t* g_rsd) {
    const uint32_t nacc = w.acc + w.dist;
    if (nacc > LZC_WINDOW) return LZC_END;
    w.acc = nacc;
    const uint32_t k = w.p - nacc;
    w.dist = g15[k];
    if (w.dist == 1u) {   // inside a run: every position back to its start carries the same 15 bytes
        const uint32_t r = g_rsd[k];
        if (w.acc + r > LZC_WINDOW) { w.acc = LZC_WINDOW; return LZC_END; }
        w.acc += r;
        w.dist = g15[k - r];
    }
    return w.dist ? LZC_GO : LZC_END;
}

}  // namespace agmvb
