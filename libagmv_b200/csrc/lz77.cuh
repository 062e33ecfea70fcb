// K4b: the LZ77 entropy coder (stream versions 3/4), AGMV_LZ77 src/agmv_encode.c:179-238.
//
// Tokens are 4 bytes: LE16 distance, u8 length, u8 "next literal". At position i the reference scans every start in
// [max(0, i-65535), i) in ascending order and keeps the first one with the strictly longest match (length capped at
// min(255, pos-i), overlap allowed, minimum length 1); then i += length + 1. Matches are long, so a frame has few
// tokens and the parse is a short serial chain: one CTA per frame walks it, and for every token all threads scan the
// window together (a thread takes every L77_THREADS-th start; the candidate is compared a word at a time against
// the token's own bytes held in shared memory). Ties go to the smallest start: each thread meets its candidates in
// ascending order, the CTA-wide reduction orders by (length desc, start asc). A candidate that reaches the cap ends
// the scan for every later start (shared atomicMin), which is what makes long runs cheap.
//
// The "next literal" of a match that ends exactly at the end of the frame's bitstream is data[pos], one byte past it:
// whatever an earlier frame left in the reference's persistent buffer (src/agmv_encode.c:218-224). The host passes,
// per frame, where that byte lives (an earlier frame of the batch, or the context's carried buffer).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace agmvb {

constexpr int L77_THREADS = 512;
constexpr int L77_MAXLEN = 255;
constexpr uint32_t L77_WINDOW = 65535;
constexpr int L77_GROUPS = (65535 + 15 + 16 + 16 * L77_THREADS - 1) / (16 * L77_THREADS);  // 16-byte groups of the window per thread

__device__ __forceinline__ uint32_t l77_load4(const uint8_t* __restrict__ p) {  // unaligned little-endian word
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}

// fs: F+1 frame starts inside bs. stale_at[f]: absolute index into bs of the byte the reference would find at data[pos]
// of frame f, or 0xFFFFFFFF -> persist[usize_f]. out: frame f's tokens at word wbase[f]. total_bits[f] = 32 * tokens.
// bs must be readable up to 8 bytes past fs[F] and 4-byte words around it (the batch buffer is padded).
__global__ void __launch_bounds__(L77_THREADS) lz77_encode_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs,
                                                             const uint32_t* __restrict__ stale_at, const uint8_t* __restrict__ persist,
                                                             const uint32_t* __restrict__ wbase, uint32_t* __restrict__ out,
                                                             uint32_t* __restrict__ total_bits) {
    __shared__ uint32_t own[L77_MAXLEN / 4 + 2];  // the token's own bytes i .. i+max-1, zero padded
    __shared__ unsigned long long wbest[L77_THREADS / 32];
    __shared__ uint32_t cap_q;                    // smallest start that reached the cap
    __shared__ uint32_t s_len;
    const uint32_t f = blockIdx.x;
    const uint8_t* __restrict__ data = bs + fs[f];
    const uint32_t n = fs[f + 1] - fs[f];
    uint32_t* __restrict__ tok = out + wbase[f];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t i = 0, T = 0;
    while (i < n) {
        const uint32_t mx = min((uint32_t)L77_MAXLEN, n - i);
        if (tid < L77_MAXLEN / 4 + 2) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t k = (uint32_t)tid * 4 + j;
                if (k < mx) w |= (uint32_t)data[i + k] << (8 * j);
            }
            own[tid] = w;
        }
        if (tid == 0) cap_q = 0xFFFFFFFFu;
        __syncthreads();
        const uint32_t lo = i > L77_WINDOW ? i - L77_WINDOW : 0u;
        const uint32_t val4 = (own[0] & 255u) * 0x01010101u;
        const uint32_t own0 = own[0];
        uint32_t blen = 0, bq = 0;
        // first-byte filter, 16 starts per load: the window is cut into 16-byte groups (aligned in memory), group g goes to
        // thread g mod L77_THREADS, all of a thread's loads are issued before the first use
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(data + lo) & ~(uintptr_t)15;
        const uint4* __restrict__ grp = reinterpret_cast<const uint4*>(a0);
        const uint32_t head = (uint32_t)(reinterpret_cast<uintptr_t>(data + lo) - a0);
        const uint32_t total = head + (i - lo);             // bytes from a0 up to (not including) position i
        const uint32_t ngroups = (total + 15u) >> 4;        // <= 4097
        uint4 v[L77_GROUPS];
#pragma unroll
        for (int k = 0; k < L77_GROUPS; k++) {
            const uint32_t g = (uint32_t)tid + (uint32_t)k * L77_THREADS;
            v[k] = g < ngroups ? grp[g] : make_uint4(~val4, ~val4, ~val4, ~val4);
        }
        bool done = false;
#pragma unroll
        for (int k = 0; k < L77_GROUPS; k++) {
            const uint32_t g = (uint32_t)tid + (uint32_t)k * L77_THREADS;
            if (g >= ngroups || done) continue;
            const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
            uint32_t m = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t t = __vcmpeq4(w[c], val4) & 0x01010101u;
                t = (t | t >> 7 | t >> 14 | t >> 21) & 15u;
                m |= t << (4 * c);
            }
            const uint32_t off0 = g << 4;                   // offset of the group's first byte from a0
            if (off0 < head) m &= 0xFFFFu << (head - off0);
            if (off0 + 16u > total) m &= 0xFFFFu >> (off0 + 16u - total);
            while (m) {
                const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
                m &= m - 1u;
                const uint32_t q = lo + (off0 + b - head);
                if (q > *reinterpret_cast<volatile uint32_t*>(&cap_q)) { done = true; break; }  // an earlier start already has the cap
                // a start can only beat this thread's best if it also matches at index blen
                if (blen > 0 && blen < mx && data[q + blen] != (uint8_t)(own[blen >> 2] >> (8 * (blen & 3)))) continue;
                uint32_t j = 0;
                uint32_t x = l77_load4(data + q) ^ own0;
                if (x) j = (uint32_t)(__ffs((int)x) - 1) >> 3;
                else {
                    j = 4;
                    while (j < mx) {
                        x = l77_load4(data + q + j) ^ own[j >> 2];
                        if (x) { j += (uint32_t)(__ffs((int)x) - 1) >> 3; break; }
                        j += 4;
                    }
                }
                j = min(j, mx);
                if (j > blen) { blen = j; bq = q; }
                if (j == mx) { atomicMin(&cap_q, q); done = true; break; }
            }
        }
        // CTA-wide: longest, then earliest
        unsigned long long key = ((unsigned long long)blen << 32) | (uint32_t)(~bq);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, d);
            key = o > key ? o : key;
        }
        if (lane == 0) wbest[warp] = key;
        __syncthreads();
        if (tid == 0) {
            unsigned long long b = 0;
#pragma unroll
            for (int w = 0; w < L77_THREADS / 32; w++) b = wbest[w] > b ? wbest[w] : b;
            const uint32_t len = (uint32_t)(b >> 32), q = ~(uint32_t)b;
            uint32_t word;
            if (len > 0) {
                uint8_t nxt;
                if (i + len < n) nxt = data[i + len];
                else nxt = stale_at[f] != 0xFFFFFFFFu ? bs[stale_at[f]] : persist[n];
                word = (i - q) | len << 16 | (uint32_t)nxt << 24;
            } else word = (own[0] & 255u) << 24;
            tok[T] = word;
            s_len = len;
        }
        __syncthreads();
        i += s_len + 1;   // a literal token advances by one as well
        T++;
    }
    if (tid == 0) total_bits[f] = T * 32u;
}

}  // namespace agmvb
