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
// So no position is ever sorted or moved: a level is one streaming pass (4 B in, 4 B out per position) plus a few
// gathers into the previous 64 KB of the same array, which sit in L1 / L2. The level-3 links come from a hashed
// "previous occurrence" table (lzchain.cuh: lzc_hashlink_k) refined by lzc_link3 below.
//
// Byte runs are the one input on which a chain walk degenerates (every position of a long run is on the chain of every
// other one). They are skipped exactly: when the walk stands on an element k whose link is 1, k-1 .. runstart(k) all carry
// the same L-gram (all one byte b) and the same byte L (b again), so if b is not the byte looked for none of them can
// end the walk, and the walk continues from the run's first position.
//
// This header is plain C++ (no CUDA types): the kernels in lzchain.cuh call these functions per thread, and
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

// link words: low 16 bits = dist, then key bytes.
//   hash level (lzc_hashlink_k): dist to the previous position with the same 3-gram hash | byte0 << 16 | byte1 << 24
//   level L >= 3:                dist_L | byte[p + L] << 16
LZC_HD uint32_t lzc_hash(uint32_t gram24, int bits) { return (gram24 * 2654435761u) >> (32 - bits); }

// level 3 from the hash chain: most recent earlier position with the same three bytes, within the window.
// d, lwh, rsd are batch arrays, p the position, rem = bytes left in p's frame from p on.
LZC_HD uint32_t lzc_link3(const uint8_t* d, const uint32_t* lwh, const uint16_t* rsd, size_t p, uint32_t rem) {
    if (rem < 3u) return 0u;
    const uint32_t w = lwh[p];
    const uint32_t key01 = w >> 16;
    const uint8_t b2 = d[p + 2];
    uint32_t dist = w & 0xFFFFu, acc = 0;
    while (dist) {
        acc += dist;
        if (acc > LZC_WINDOW) return 0u;
        const size_t k = p - acc;
        const uint32_t wk = lwh[k];
        const uint32_t g01 = wk >> 16;
        if (g01 == key01 && d[k + 2] == b2) return acc;
        dist = wk & 0xFFFFu;
        // run skip: k reads b,b,b and so does k-1 (its link is 1): every position back to the run start has the same
        // gram, hence the same hash, and is the next chain element; none of them is p's gram (k was not)
        if (dist == 1u && (g01 & 0xFFu) == (g01 >> 8) && d[k + 2] == (uint8_t)g01) {
            const uint32_t r = rsd[k];
            if (r) {
                if (acc + r > LZC_WINDOW) return 0u;
                acc += r;
                dist = lwh[p - acc] & 0xFFFFu;
            }
        }
    }
    return 0u;
}

// one level: dist_{L+1}[p] from the level-L links. Returns the new distance (0 = no match of length L+1).
// *rec is set to L << 28 | LZC_RESOLVED | offset when p had a match of length L and has none of length L+1 (its final
// answer: offset = distance to the EARLIEST level-L occurrence inside the window), and left untouched otherwise.
LZC_HD uint32_t lzc_level(const uint8_t* d, const uint32_t* lw, const uint16_t* rsd, size_t p, uint32_t rem, uint32_t L, uint32_t* rec) {
    const uint32_t w = lw[p];
    uint32_t dist = w & 0xFFFFu;
    if (!dist) return 0u;
    const uint32_t c = (w >> 16) & 0xFFu;      // byte L of p
    const bool can_extend = L + 1u <= rem;     // the reference caps the match at the bytes left in the frame (src/agmv_encode.c:121-123)
    const uint32_t b = d[p];
    uint32_t acc = 0, last = 0;
    while (dist) {
        acc += dist;
        if (acc > LZC_WINDOW) break;
        const uint32_t wk = lw[p - acc];
        if (can_extend && ((wk >> 16) & 0xFFu) == c) return acc;
        last = acc;
        dist = wk & 0xFFFFu;
        if (dist == 1u && (!can_extend || c != b)) {
            // the element stands inside a run of byte b = p's own first byte (same L-gram); everything back to the run
            // start is on the chain and carries byte L == b != c
            const uint32_t r = rsd[p - acc];
            if (acc + r > LZC_WINDOW) { last = LZC_WINDOW; break; }   // the window ends inside the run: p - 65535 is on the chain
            acc += r;
            last = acc;
            dist = lw[p - acc] & 0xFFFFu;
        }
    }
    *rec = L << 28 | LZC_RESOLVED | last;
    return 0u;
}

// positions whose match reaches 15 bytes: offset of the earliest occurrence of the 15-gram inside the window
LZC_HD uint32_t lzc_chain_end(const uint32_t* lw15, const uint16_t* rsd, size_t p) {
    uint32_t dist = lw15[p] & 0xFFFFu, acc = 0, last = 0;
    while (dist) {
        acc += dist;
        if (acc > LZC_WINDOW) break;
        last = acc;
        dist = lw15[p - acc] & 0xFFFFu;
        if (dist == 1u) {
            const uint32_t r = rsd[p - acc];
            if (acc + r > LZC_WINDOW) { last = LZC_WINDOW; break; }
            acc += r;
            last = acc;
            dist = lw15[p - acc] & 0xFFFFu;
        }
    }
    return last;
}

}  // namespace agmvb
