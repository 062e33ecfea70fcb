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
// "previous occurrence" table (lzchain.cuh: lzc_hashlink_k) refined by LzcLink3Walk below.
//
// Byte runs are the one input on which a chain walk degenerates (every position of a long run is on the chain of every
// other one). They are skipped exactly: when the walk stands on an element k whose link is 1, k-1 .. runstart(k) all carry
// the same L-gram (all one byte b) and the same byte L (b again), so if b is not the byte looked for none of them can
// end the walk, and the walk continues from the run's first position.
//
// This header is plain C++ (no CUDA types): the kernels in lzchain.cuh drive these state machines one hop at a time, and
// tests/lzchain_host_check.cpp compiles them for the host to check the logic against a brute-force search.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define LZC_HD __host__ __device__ __forceinline__
#define LZC_HDM __host__ __device__ __forceinline__
#else
#define LZC_HD static inline
#define LZC_HDM inline
#endif

namespace agmvb {

constexpr uint32_t LZC_WINDOW = 65535u;     // src/agmv_encode.c:102
constexpr uint32_t LZC_DEAD = 1u << 29;     // link word of a position whose match is final (see below)

// Link words: low 16 bits = dist, then key bits.
//   hash level (lzc_hashlink_k): dist to the previous position with the same 3-gram hash | byte0 << 16 | byte1 << 24
//   level L >= 3, position still growing ("live", dist_L != 0):
//                  dist_L | byte[p+L] << 16 | cap << 24 | (byte[p+L] != byte[p+L-1]) << 28,
//                  cap = min(15, bytes left in the frame from p on): the reference caps a match there (src/agmv_encode.c:121-123)
//   level L >= 3, position whose match is final ("dead", dist_L == 0): the word carries the RESULT, so that no level writes
//                  anything but its own dense output array:
//                  offset | byte[p+L] << 16 | length << 24 | LZC_DEAD   (length 0: no match, a literal)
//                  A dead word is copied from level to level (only its byte field changes); walkers that land on it read its
//                  byte like any other element's and its link as 0.
LZC_HD uint32_t lzc_hash(uint32_t gram24, int bits) { return (gram24 * 2654435761u) >> (32 - bits); }
LZC_HD uint32_t lzc_word(uint32_t dist, uint32_t byte_l, uint32_t byte_before, uint32_t cap) {
    return dist | byte_l << 16 | cap << 24 | (byte_l != byte_before ? 1u << 28 : 0u);
}
LZC_HD uint32_t lzc_dead(uint32_t len, uint32_t off, uint32_t byte_l) { return off | byte_l << 16 | len << 24 | LZC_DEAD; }
LZC_HD uint32_t lzc_link(uint32_t w) { return (w & LZC_DEAD) ? 0u : (w & 0xFFFFu); }   // dist_L of a level word
LZC_HD uint32_t lzc_carry(uint32_t w, uint32_t byte_l) { return (w & 0xFF00FFFFu) | byte_l << 16; }   // a dead word at the next level

enum { LZC_GO = 0, LZC_FOUND = 1, LZC_END = 2 };

// level 3 from the hash chain: most recent earlier position with the same three bytes, within the window.
// Usage: if (start(...)) while ((r = hop(...)) == LZC_GO); r == LZC_FOUND: acc = dist_3.
struct LzcLink3Walk {
    uint32_t p, acc, dist, key;   // key = byte0 | byte1 << 8 | byte2 << 16
    LZC_HDM bool start(uint32_t p_, uint32_t w, uint32_t byte2, uint32_t cap) {
        p = p_; acc = 0; dist = w & 0xFFFFu; key = (w >> 16) | byte2 << 16;
        return cap >= 3u && dist != 0u;
    }
    LZC_HDM int hop(const uint8_t* d, const uint32_t* lwh, const uint16_t* rsd) {
        acc += dist;
        if (acc > LZC_WINDOW) return LZC_END;
        const uint32_t k = p - acc, wk = lwh[k], g01 = wk >> 16;
        if (g01 == (key & 0xFFFFu) && d[k + 2] == (key >> 16)) return LZC_FOUND;
        dist = wk & 0xFFFFu;
        // run skip: k reads b,b,b and so does k-1 (its link is 1): every position back to the run start has the same
        // gram, hence the same hash, and is the next chain element; none of them is p's gram (k was not)
        if (dist == 1u && (g01 & 0xFFu) == (g01 >> 8) && d[k + 2] == (g01 & 0xFFu)) {
            const uint32_t r = rsd[k];
            if (r) {
                if (acc + r > LZC_WINDOW) return LZC_END;
                acc += r;
                dist = lwh[p - acc] & 0xFFFFu;
            }
        }
        return dist ? LZC_GO : LZC_END;
    }
};

// one level: dist_{L+1}[p] from the level-L links.
// r == LZC_FOUND: acc = dist_{L+1}. r == LZC_END: no match of length L+1; `last` = distance to the EARLIEST level-L
// occurrence inside the window, i.e. the offset of p's final (length L) match.
struct LzcLevelWalk {
    uint32_t p, acc, last, dist, c;
    bool ext, neq;
    LZC_HDM bool start(uint32_t p_, uint32_t w, uint32_t L) {
        p = p_; acc = 0; last = 0; dist = lzc_link(w); c = (w >> 16) & 0xFFu;
        ext = L + 1u <= ((w >> 24) & 0xFu);
        neq = (w >> 28) & 1u;
        return dist != 0u;
    }
    LZC_HDM int hop(const uint32_t* lw, const uint16_t* rsd) {
        acc += dist;
        if (acc > LZC_WINDOW) return LZC_END;
        const uint32_t wk = lw[p - acc];
        if (ext && ((wk >> 16) & 0xFFu) == c) return LZC_FOUND;
        last = acc;
        dist = lzc_link(wk);
        if (dist == 1u && (!ext || neq)) {
            // the element stands inside a run of one byte b (its L-gram, hence p's, is all b, and neq says p's byte L is
            // not b): everything back to the run start is on the chain and carries byte L == b
            const uint32_t r = rsd[p - acc];
            if (acc + r > LZC_WINDOW) { last = LZC_WINDOW; return LZC_END; }   // the window ends inside the run: p - 65535 is on the chain
            acc += r;
            last = acc;
            dist = lzc_link(lw[p - acc]);
        }
        return dist ? LZC_GO : LZC_END;
    }
};

// positions whose match reaches 15 bytes: `last` ends as the offset of the earliest occurrence of the 15-gram inside the
// window (the end of the level-15 chain). Usage: start(); while (hop(...) == LZC_GO);
struct LzcEndWalk {
    uint32_t p, acc, last, dist;
    LZC_HDM bool start(uint32_t p_, uint32_t w) { p = p_; acc = 0; last = 0; dist = lzc_link(w); return dist != 0u; }
    LZC_HDM int hop(const uint32_t* lw15, const uint16_t* rsd) {
        acc += dist;
        if (acc > LZC_WINDOW) return LZC_END;
        last = acc;
        dist = lzc_link(lw15[p - acc]);
        if (dist == 1u) {   // inside a run: every position back to its start carries the same 15 bytes
            const uint32_t r = rsd[p - acc];
            if (acc + r > LZC_WINDOW) { last = LZC_WINDOW; return LZC_END; }
            acc += r;
            last = acc;
            dist = lzc_link(lw15[p - acc]);
        }
        return dist ? LZC_GO : LZC_END;
    }
};
LZC_HD uint32_t lzc_chain_end(const uint32_t* lw15, const uint16_t* rsd, uint32_t p) {
    LzcEndWalk w;
    if (w.start(p, lw15[p])) while (w.hop(lw15, rsd) == LZC_GO) {}
    return w.last;
}

}  // namespace agmvb
