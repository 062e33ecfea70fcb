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
#include "radix.cuh"

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

// match length of cand[0..] against the token's own bytes (own: words in shared memory, zero padded, at least
// L77_OWN_WORDS long), capped at mx. Eight words per step, all loads of a step in flight together: a 255-byte match takes
// 8 dependent steps instead of 64.
constexpr int L77_OWN_WORDS = L77_MAXLEN / 4 + 10;
__device__ __forceinline__ uint32_t l77_lcp(const uint8_t* __restrict__ cand, const uint32_t* own, uint32_t mx) {
    uint32_t j = 0;
    while (j < mx) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(cand + j);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3) * 8;
        uint32_t v[9];
#pragma unroll
        for (int k = 0; k < 9; k++) v[k] = w[k];
        uint32_t first = 32;   // bytes matched inside this step
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            const uint32_t x = __funnelshift_r(v[k], v[k + 1], sh) ^ own[(j >> 2) + k];
            if (x) first = 4u * k + ((uint32_t)(__ffs((int)x) - 1) >> 3);
        }
        j += first;
        if (first < 32) break;
    }
    return min(j, mx);
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

// =====================================================================================================
// Faster variant: candidates from sorted bucket lists instead of a window scan.
//
// Two stable counting-sort passes (radix.cuh) order the positions of the whole batch by (byte 1, byte 0, position):
// every 2-byte prefix owns a contiguous, position-sorted slice of A2 (bucket table b2), every first byte a slice of A1.
// For a token at position P the starts that match at least two bytes are the entries of P's bucket that lie in
// [max(frame start, P-65535), P): two cooperative searches bound them, and only they are extended (in ascending order,
// a CTA-wide chunk at a time, so "first longest" is "first chunk that raises the maximum, smallest start inside it";
// a candidate is extended only if it also matches at index best). If there is none, the answer is length 1 at the
// first entry of the first byte's slice inside the window, or a literal.
// =====================================================================================================
constexpr int T77 = 128;
constexpr int L77_U = 4;      // starts per thread and chunk

struct L77Byte0 { const uint8_t* bs; __device__ uint32_t operator()(uint32_t i) const { return bs[i]; } };
struct L77Byte1 { const uint8_t* bs; const uint32_t* a1; __device__ uint32_t operator()(uint32_t i) const { return bs[a1[i] + 1]; } };
struct L77Move0 { uint32_t* out; __device__ void operator()(uint32_t s, uint32_t d) const { out[d] = s; } };
struct L77Move1 { const uint32_t* in; uint32_t* out; __device__ void operator()(uint32_t s, uint32_t d) const { out[d] = in[s]; } };

// slice boundaries of the 2-byte buckets in A2 (tables zeroed by the caller: start == end == 0 means empty)
__global__ void __launch_bounds__(256) l77_buckets_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ a2, uint32_t n,
                                                     uint32_t* __restrict__ b2s, uint32_t* __restrict__ b2e, uint32_t* __restrict__ inv2) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = a2[i];
    inv2[p] = i;   // where position p sits in A2: everything before it in its bucket starts earlier
    const uint32_t k = (uint32_t)bs[p] | (uint32_t)bs[p + 1] << 8;
    uint32_t kp = 0x10000u, kn = 0x10000u;
    if (i > 0) { const uint32_t q = a2[i - 1]; kp = (uint32_t)bs[q] | (uint32_t)bs[q + 1] << 8; }
    if (i + 1 < n) { const uint32_t q = a2[i + 1]; kn = (uint32_t)bs[q] | (uint32_t)bs[q + 1] << 8; }
    if (k != kp) b2s[k] = i;
    if (k != kn) b2e[k] = i + 1;
}
// first-byte slices of A1 from the scanned tile histogram of the first pass (entry [d * ntiles] = first index of byte d)
__global__ void l77_b1_k(const uint32_t* __restrict__ tile_hist, uint32_t ntiles, uint32_t n, uint32_t* __restrict__ b1) {
    const uint32_t d = threadIdx.x;
    if (d < 256) b1[d] = tile_hist[d * ntiles];
    if (d == 0) b1[256] = n;
}

// first index in [lo, hi) of the ascending array a with a[idx] >= target (hi if none); all T77 threads call it together
__device__ __forceinline__ uint32_t l77_coop_lower_bound(const uint32_t* __restrict__ a, uint32_t lo, uint32_t hi, uint32_t target,
                                                         uint32_t* wfirst) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    while (lo < hi) {
        const uint32_t m = hi - lo;
        const uint32_t step = (m + T77 - 1) / T77;
        const uint32_t last = lo + min((uint32_t)(tid + 1) * step, m) - 1u;     // probe: the last element of this thread's piece
        const bool have = (uint32_t)tid * step < m;
        const bool pred = have && a[last] >= target;
        const unsigned bal = __ballot_sync(0xffffffffu, pred);
        if (lane == 0) wfirst[warp] = bal ? (uint32_t)(warp * 32 + __ffs(bal) - 1) : 0xFFFFFFFFu;
        __syncthreads();
        uint32_t tf = 0xFFFFFFFFu;
#pragma unroll
        for (int w = 0; w < T77 / 32; w++) tf = min(tf, wfirst[w]);
        __syncthreads();
        if (tf == 0xFFFFFFFFu) return hi;
        const uint32_t nlo = lo + tf * step, nhi = lo + min((tf + 1) * step, m);   // a[nhi - 1] >= target
        if (step == 1) return nlo;
        lo = nlo;
        hi = nhi - 1;        // the answer is in [nlo, nhi - 1]; if nothing in [nlo, nhi - 1) qualifies it is nhi - 1
        if (lo == hi) return lo;
        // search [lo, hi) and fall back to hi (= nhi - 1, known to qualify)
    }
    return lo;
}

__global__ void __launch_bounds__(T77) lz77_bucket_encode_k(const uint8_t* __restrict__ bs, const uint32_t* __restrict__ fs,
                                                            const uint32_t* __restrict__ a1, const uint32_t* __restrict__ a2,
                                                            const uint32_t* __restrict__ b1, const uint32_t* __restrict__ b2s,
                                                            const uint32_t* __restrict__ b2e, const uint32_t* __restrict__ inv2,
                                                            const uint32_t* __restrict__ stale_at,
                                                            const uint8_t* __restrict__ persist, const uint32_t* __restrict__ wbase,
                                                            uint32_t* __restrict__ out, uint32_t* __restrict__ total_bits) {
    __shared__ uint32_t own[L77_OWN_WORDS];
    __shared__ unsigned long long wbest[T77 / 32];
    __shared__ uint32_t wfirst[T77 / 32];
    const uint32_t f = blockIdx.x;
    const uint32_t F0 = fs[f], n = fs[f + 1] - F0;
    const uint8_t* __restrict__ data = bs + F0;
    uint32_t* __restrict__ tok = out + wbase[f];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t i = 0, T = 0;
    while (i < n) {
        const uint32_t mx = min((uint32_t)L77_MAXLEN, n - i);
        if (tid < L77_OWN_WORDS) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t k = (uint32_t)tid * 4 + j;
                if (k < mx) w |= (uint32_t)data[i + k] << (8 * j);
            }
            own[tid] = w;
        }
        __syncthreads();
        const uint32_t P = F0 + i;
        const uint32_t LO = (i > L77_WINDOW) ? P - L77_WINDOW : F0;
        const uint32_t own0 = own[0];
        uint32_t blen = 0, bq = 0;   // CTA-uniform
        if (mx >= 2) {
            const uint32_t key = own0 & 0xFFFFu;
            const uint32_t s = b2s[key];
            const uint32_t c1 = inv2[P];            // P's own slot: the starts before it in the bucket are exactly the earlier ones
            // first start inside the window: walk back from c1 in strides of 32 (one probe per thread covers 4096 starts),
            // then resolve inside the stride; wider windows of candidates fall back to the full search
            uint32_t c0;
            {
                const uint32_t back = (uint32_t)(tid + 1) * 32u;
                const bool inb = c1 >= s + back;                         // slot c1 - back exists in the bucket
                const bool below = inb && a2[c1 - back] < LO;            // ... and lies before the window
                const unsigned bal = __ballot_sync(0xffffffffu, below);
                if (lane == 0) wfirst[warp] = bal ? (uint32_t)(warp * 32 + __ffs(bal) - 1) : 0xFFFFFFFFu;
                __syncthreads();
                uint32_t tf = 0xFFFFFFFFu;
#pragma unroll
                for (int w = 0; w < T77 / 32; w++) tf = min(tf, wfirst[w]);
                __syncthreads();
                uint32_t lo, hi;   // answer in [lo, hi]: a2[lo - 1] < LO (or lo == s), a2[hi] >= LO (or hi == c1)
                if (tf != 0xFFFFFFFFu) { lo = c1 - (tf + 1u) * 32u + 1u; hi = c1 - tf * 32u; }
                else if (c1 - s <= (uint32_t)T77 * 32u) { lo = s; hi = c1 - min(c1 - s, ((c1 - s) / 32u) * 32u); }
                else { lo = s; hi = c1 - (uint32_t)T77 * 32u; }
                c0 = l77_coop_lower_bound(a2, lo, hi, LO, wfirst);
            }
            // A start can only be the answer if its match is at least as long as ANY other start's. The most recent starts
            // tend to match longest: measure them first (length only) and let the ordered scan discard, with one byte load,
            // every start that does not reach that length.
            uint32_t floor_len = 2;
            if (c1 - c0 > (uint32_t)(T77 * L77_U)) {
                const uint32_t idx = c1 - T77 + tid;
                uint32_t j = l77_lcp(bs + a2[idx], own, mx);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) j = max(j, __shfl_xor_sync(0xffffffffu, j, d));
                if (lane == 0) wfirst[warp] = j;
                __syncthreads();
#pragma unroll
                for (int w = 0; w < T77 / 32; w++) floor_len = max(floor_len, wfirst[w]);
                __syncthreads();
            }
            for (uint32_t base = c0; base < c1; base += T77 * L77_U) {
                uint32_t q[L77_U], jj[L77_U];
                bool go[L77_U];
                // a start matters if it reaches floor_len and is strictly longer than the best of the earlier starts
                const uint32_t need = max(floor_len, blen + 1u);   // minimum useful length (blen + 1 > mx cannot happen: the loop ends at mx)
#pragma unroll
                for (int u = 0; u < L77_U; u++) {
                    const uint32_t idx = base + (uint32_t)u * T77 + tid;
                    go[u] = idx < c1;
                    q[u] = go[u] ? a2[idx] : 0u;
                }
#pragma unroll
                for (int u = 0; u < L77_U; u++)
                    if (go[u] && need > 2) go[u] = bs[q[u] + need - 1] == (uint8_t)(own[(need - 1) >> 2] >> (8 * ((need - 1) & 3)));
                unsigned long long k2 = 0;
#pragma unroll
                for (int u = 0; u < L77_U; u++) {
                    jj[u] = 0;
                    if (go[u]) {
                        jj[u] = l77_lcp(bs + q[u], own, mx);
                    }
                    const unsigned long long ku = ((unsigned long long)jj[u] << 32) | (uint32_t)(~q[u]);
                    k2 = ku > k2 ? ku : k2;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k2, d);
                    k2 = o > k2 ? o : k2;
                }
                if (lane == 0) wbest[warp] = k2;
                __syncthreads();
                unsigned long long bb = 0;
#pragma unroll
                for (int w = 0; w < T77 / 32; w++) bb = wbest[w] > bb ? wbest[w] : bb;
                __syncthreads();
                if ((uint32_t)(bb >> 32) > blen) { blen = (uint32_t)(bb >> 32); bq = ~(uint32_t)bb; }
                if (blen == mx) break;
            }
        }
        if (blen == 0) {   // no start matches two bytes (or only one byte is left): the earliest start with the same first byte
            const uint32_t v0 = own0 & 255u;
            const uint32_t s = b1[v0], e = b1[v0 + 1];
            const uint32_t c = l77_coop_lower_bound(a1, s, e, LO, wfirst);
            if (c < e) {
                const uint32_t q = a1[c];
                if (q < P) { blen = 1; bq = q; }
            }
        }
        if (tid == 0) {
            uint32_t word;
            if (blen > 0) {
                uint8_t nxt;
                if (i + blen < n) nxt = data[i + blen];
                else nxt = stale_at[f] != 0xFFFFFFFFu ? bs[stale_at[f]] : persist[n];
                word = (P - bq) | blen << 16 | (uint32_t)nxt << 24;
            } else word = (own0 & 255u) << 24;
            tok[T] = word;
        }
        __syncthreads();   // own[] is rewritten by the next token
        i += blen + 1;
        T++;
    }
    if (tid == 0) total_bits[f] = T * 32u;
}

}  // namespace agmvb
