// Encode-side kernels K0a (histogram), K0b (palette), K1+K2 (interpolate +
// quantise), K3 (4x4 block classification + bitstream assembly).
#pragma once
#include "common.cuh"
#include "radix.cuh"
#include "scan.cuh"

namespace agmvb {

// ---------------------------------------------------------------------------
// K0a: quantised-colour histogram over source frames.
// Reference: src/agmv_encode.c:2389-2394 with AGMV_QuantizeColor
// (src/agmv_utils.c:695-742). Bin max_clr (an out-of-bounds write in the
// reference, never selectable) is dropped. Counters are 64-bit: a 4K x 8000
// frame job has 6.6e10 pixels. Lanes that hit the same bin are merged with
// __match_any_sync before the L2 atomic (flat synthetic/cartoon content would
// otherwise serialise 32 atomics on one address).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void hist_add(unsigned long long* hist, uint32_t q, bool valid) {
    uint32_t key = valid ? q : 0x80000000u + lane_id();
    unsigned peers = __match_any_sync(0xffffffffu, key);
    if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&hist[q], (unsigned long long)__popc(peers));
}

// Four pixels per thread (one 16-byte load), so a warp sees 128 consecutive pixels. Equal quantised colours are merged by
// RUNS of consecutive pixels rather than by MATCH rounds (profiles/r01: four MATCH per thread bound the kernel at 21 % of
// HBM): a pixel whose bin differs from its predecessor's heads a run, the four head bitmaps (one ballot per pixel slot)
// give every head its run length with a few bit operations, and the head issues ONE atomic for the whole run. Flat
// content costs an atomic per run of up to 128 pixels, a gradient one per two or three pixels; equal colours that are not
// adjacent simply take separate atomics. (A pixel in bin max_clr is dropped, as above; it still ends a run.) The first
// version took four ballots and a 4 x 4 search per head: issue-bound at 76 % with 44 instructions per pixel
// (profiles/r02_ncu_summary.txt).
__global__ void __launch_bounds__(256) hist_vec4_k(const uint4* __restrict__ px4, uint64_t ngroups, int quality, uint32_t mc,
                                                   unsigned long long* __restrict__ hist) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t lane = lane_id();
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < ngroups; base += stride) {
        const uint64_t g = base + threadIdx.x;
        const bool valid = g < ngroups;
        const uint4 v = valid ? __ldg(px4 + g) : make_uint4(0, 0, 0, 0);
        uint32_t q[4];
        q[0] = valid ? quantize_color(v.x, quality) : 0xFFFFFFFFu;
        q[1] = valid ? quantize_color(v.y, quality) : 0xFFFFFFFFu;
        q[2] = valid ? quantize_color(v.z, quality) : 0xFFFFFFFFu;
        q[3] = valid ? quantize_color(v.w, quality) : 0xFFFFFFFFu;
        const uint32_t left = __shfl_up_sync(0xffffffffu, q[3], 1);
        const bool hd[4] = {lane == 0 || q[0] != left, q[1] != q[0], q[2] != q[1], q[3] != q[2]};
        // where the run of this lane's LAST head ends: at the first head of the next lane that has one (one ballot, one
        // shuffle); every earlier head of the lane ends at the next head of the same lane (registers only)
        const unsigned anyb = __ballot_sync(0xffffffffu, hd[0] | hd[1] | hd[2] | hd[3]);
        const uint32_t first = hd[0] ? 0u : (hd[1] ? 1u : (hd[2] ? 2u : (hd[3] ? 3u : 4u)));
        const unsigned above = lane < 31u ? anyb >> (lane + 1u) : 0u;
        const uint32_t nl = above ? lane + (uint32_t)__ffs(above) : 32u;
        const uint32_t nfirst = __shfl_sync(0xffffffffu, first, (int)(nl & 31u));
        uint32_t nx = nl < 32u ? 4u * nl + nfirst : 128u;   // pixel position (0..128) of the next head after this lane
#pragma unroll
        for (int k = 3; k >= 0; k--) {
            const uint32_t at = 4u * lane + (uint32_t)k;
            if (hd[k]) {
                if (q[k] < mc) atomicAdd(&hist[q[k]], (unsigned long long)(nx - at));   // (the dropped bin / padding still ends a run)
                nx = at;
            }
        }
    }
}

__global__ void __launch_bounds__(256) hist_scalar_k(const uint32_t* __restrict__ px, uint64_t n, int quality, uint32_t mc,
                                                     unsigned long long* __restrict__ hist) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += stride) {
        uint64_t i = base + threadIdx.x;
        bool valid = i < n;
        uint32_t q = quantize_color(valid ? px[i] : 0u, quality);
        hist_add(hist, q, valid && q < mc);
    }
}

// ---------------------------------------------------------------------------
// K0b: palette construction.
// AGMV_BubbleSort (src/agmv_utils.c:995-1010) is a stable ascending sort by
// count carrying the colour index: identical to sorting the 64-bit keys
// count << 19 | index. The greedy pick (src/agmv_encode.c:2572-2625) walks the
// sorted list from the top and rejects a colour that lies within a per-channel
// box of ANY of the 512 palette slots - including still-empty slots, which hold
// 0, i.e. quantised black. Because the box test is symmetric, "is candidate c
// rejected" == "c lies in the box of black or of an already picked colour", so
// a bitmap over the quantised colour cube (<= 2^19 bits, in shared memory)
// that gets the box of every pick OR-ed in answers it in O(1).
// ---------------------------------------------------------------------------
__global__ void pal_keys_k(const unsigned long long* __restrict__ hist, uint32_t mc, unsigned long long* __restrict__ keys) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < mc) keys[i] = (hist[i] << 19) | i;
}

struct KeyDigit {
    const unsigned long long* k;
    uint32_t shift;
    __device__ uint32_t operator()(uint32_t i) const { return (uint32_t)(k[i] >> shift) & 255u; }
};
struct KeyMove {
    const unsigned long long* in;
    unsigned long long* out;
    __device__ void operator()(uint32_t s, uint32_t d) const { out[d] = in[s]; }
};

struct QBits { int rs, gs, rmax, gmax, bmax, tr, tg, tb; };
__device__ __forceinline__ QBits qbits(int q) {
    QBits b;
    if (q == Q_MID) { b.rs = 12; b.gs = 6; b.rmax = 31; b.gmax = 63; b.bmax = 63; }
    else if (q == Q_LOW) { b.rs = 11; b.gs = 5; b.rmax = 31; b.gmax = 63; b.bmax = 31; }
    else { b.rs = 13; b.gs = 7; b.rmax = 63; b.gmax = 63; b.bmax = 127; }
    if (q == Q_HIGH) { b.tr = 2; b.tg = 2; b.tb = 3; } else { b.tr = 1; b.tg = 1; b.tb = 1; }  // src/agmv_encode.c:2605-2614
    return b;
}
// src/agmv_utils.c:744-783
__device__ __forceinline__ uint32_t dequantize(uint32_t c, int q) {
    QBits b = qbits(q);
    uint32_t r = (c >> b.rs) & b.rmax, g = (c >> b.gs) & b.gmax, bl = c & b.bmax;
    if (q == Q_MID) { r <<= 3; g <<= 2; bl <<= 2; }
    else if (q == Q_LOW) { r <<= 3; g <<= 2; bl <<= 3; }
    else { r <<= 2; g <<= 2; bl <<= 1; }
    return r << 16 | g << 8 | bl;
}

__device__ __forceinline__ void pal_block_box(uint32_t* blocked, uint32_t c, const QBits& b) {
    int r = (c >> b.rs) & b.rmax, g = (c >> b.gs) & b.gmax, bl = c & b.bmax;
    int nr = 2 * b.tr + 1, ng = 2 * b.tg + 1, nb = 2 * b.tb + 1;
    for (int k = lane_id(); k < nr * ng * nb; k += 32) {
        int rr = r + k / (ng * nb) - b.tr, gg = g + (k / nb) % ng - b.tg, bb = bl + k % nb - b.tb;
        if (rr < 0 || rr > b.rmax || gg < 0 || gg > b.gmax || bb < 0 || bb > b.bmax) continue;
        uint32_t idx = (uint32_t)rr << b.rs | (uint32_t)gg << b.gs | (uint32_t)bb;
        atomicOr(&blocked[idx >> 5], 1u << (idx & 31));
    }
}

// one warp; dynamic smem = (mc + 1) / 8 bytes
__global__ void __launch_bounds__(32) pal_pick_k(const unsigned long long* __restrict__ sorted, uint32_t mc, int quality, int dual,
                                                 uint32_t* __restrict__ pal_out /* pal0[256] then pal1[256] */) {
    extern __shared__ uint32_t blocked[];
    __shared__ uint32_t pal[512];
    const int lane = lane_id();
    const QBits qb = qbits(quality);
    for (uint32_t k = lane; k < (mc + 1) / 32; k += 32) blocked[k] = 0;
    for (int k = lane; k < 512; k += 32) { pal[k] = 0; pal_out[k] = 0; }
    __syncwarp();
    pal_block_box(blocked, 0u, qb);  // empty palette slots compare as quantised black
    __syncwarp();
    int count = 0;
    for (int64_t n0 = (int64_t)mc - 1; n0 >= 1 && count < 512; n0 -= 32) {
        int64_t n = n0 - lane;
        bool alive = n >= 1;
        uint32_t clr = alive ? (uint32_t)(sorted[n] & 0x7FFFFu) : 0u;
        while (true) {
            bool ok = alive && !((blocked[clr >> 5] >> (clr & 31)) & 1u);
            unsigned m = __ballot_sync(0xffffffffu, ok);
            if (!m) break;
            int first = __ffs(m) - 1;
            uint32_t c = __shfl_sync(0xffffffffu, clr, first);
            if (lane == 0) pal[count] = c;
            count++;
            pal_block_box(blocked, c, qb);
            __syncwarp();
            alive = alive && lane > first;
            if (count >= 512) break;
        }
    }
    __syncwarp();
    // split + de-quantise, src/agmv_encode.c:2627-2656
    for (int n = lane; n < 512; n += 32) {
        uint32_t inv = dequantize(pal[n], quality);
        if (dual) {
            if (n < 126) pal_out[n] = inv;
            else if (n <= 252) pal_out[256 + n - 126] = inv;
            else if (n <= 381) pal_out[n - 126] = inv;
            else if (n - 255 < 256) pal_out[256 + n - 255] = inv;
        } else if (n < 256) pal_out[n] = inv;
    }
}

// ---------------------------------------------------------------------------
// K1+K2: frame interpolation fused with nearest-palette-entry quantisation.
// AGMV_InterpFrame (src/agmv_utils.c:949-969): c = c1 + ((c2 - c1) >> 1) per
// channel. AGMV_FindNearestEntry / AGMV_FindNearestColor
// (src/agmv_utils.c:785-895): argmin of the squared RGB distance, lowest index
// on ties, palette 0 on a tie between palettes - i.e. the minimum of
// dist * 512 + (pal_num * 256 + index). The argmin is a pure function of the
// 24-bit colour, so it is memoised in a 2^24-entry table (32 MB, L2 resident):
// a warp that misses evaluates the 512 distances cooperatively (16 per lane,
// palette in shared memory) and publishes the entry; every later pixel of that
// colour in the whole video is one 2-byte gather.
// ---------------------------------------------------------------------------
struct SrcPair { const uint32_t* a; const uint32_t* b; };  // b == nullptr: no interpolation

__device__ __forceinline__ uint32_t interp_px(uint32_t c1, uint32_t c2) {
    int r1 = (c1 >> 16) & 255, g1 = (c1 >> 8) & 255, b1 = c1 & 255;
    int r2 = (c2 >> 16) & 255, g2 = (c2 >> 8) & 255, b2 = c2 & 255;
    int r = r1 + ((r2 - r1) >> 1), g = g1 + ((g2 - g1) >> 1), b = b1 + ((b2 - b1) >> 1);
    return (uint32_t)(r << 16 | g << 8 | b);
}

__device__ __forceinline__ uint32_t lut_entry(uint32_t c, bool valid, uint16_t* __restrict__ lut, const uint32_t* spal, int npal) {
    uint32_t e = valid ? lut[c] : 0u;
    bool miss = valid && e == LUT_EMPTY;
    unsigned m = __ballot_sync(0xffffffffu, miss);
    while (m) {
        int leader = __ffs(m) - 1;
        uint32_t cc = __shfl_sync(0xffffffffu, c, leader);
        // |c - p|^2 = c.c + p.p - 2 c.p on the packed bytes (top byte 0 on both sides): exact in 32 bits, two dp4a per entry
        const uint32_t cdot = __dp4a(cc, cc, 0u);
        uint32_t best = 0xFFFFFFFFu;
        for (int j = lane_id(); j < npal; j += 32) {
            const uint32_t p = spal[j];
            const uint32_t d = cdot + __dp4a(p, p, 0u) - 2u * __dp4a(cc, p, 0u);
            const uint32_t key = d * 512u + (uint32_t)j;
            best = key < best ? key : best;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            uint32_t o = __shfl_xor_sync(0xffffffffu, best, d);
            best = o < best ? o : best;
        }
        uint32_t ent = best & 511u;
        if (miss && c == cc) { e = ent; miss = false; }
        if ((int)lane_id() == leader) lut[cc] = (uint16_t)ent;
        m = __ballot_sync(0xffffffffu, miss);
    }
    return e;
}

// The whole table at once: 2^24 colours x 512 entries is 8.6 G distances - about a millisecond of dp4a on this GPU, less
// than what the cooperative miss path costs while a long sequence warms the table up (the first 1070-frame batch of the
// bench configuration quantised at 7.6 us per frame cold against 4.8 us warm, profiles/r02_ncu_summary.txt). Used when a
// call encodes enough pixels to pay for it (api.cu); the quantisers then never meet an empty entry.
// Ordering: key = d * 512 + j with d = c.c + p.p - 2 c.p; c.c is the same for every j, so the minimum of
// (p.p * 512 + j) - 1024 c.p (signed) picks the same entry, and j = key & 511.
// grid 2^24 / 1024 blocks of 256 threads, four consecutive colours per thread.
__global__ void __launch_bounds__(256) lut_fill_k(const uint32_t* __restrict__ pal, int npal, uint2* __restrict__ lut4) {
    __shared__ uint2 sp[512];
    for (int k = threadIdx.x; k < npal; k += blockDim.x) {
        const uint32_t p = pal[k];
        sp[k] = make_uint2(p, __dp4a(p, p, 0u) * 512u + (uint32_t)k);
    }
    __syncthreads();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c0 = t * 4u;
    int b0 = 0x7fffffff, b1 = 0x7fffffff, b2 = 0x7fffffff, b3 = 0x7fffffff;
#pragma unroll 8
    for (int j = 0; j < npal; j++) {
        const uint2 e = sp[j];
        b0 = min(b0, (int)e.y - 1024 * (int)__dp4a(c0, e.x, 0u));
        b1 = min(b1, (int)e.y - 1024 * (int)__dp4a(c0 + 1u, e.x, 0u));
        b2 = min(b2, (int)e.y - 1024 * (int)__dp4a(c0 + 2u, e.x, 0u));
        b3 = min(b3, (int)e.y - 1024 * (int)__dp4a(c0 + 3u, e.x, 0u));
    }
    lut4[t] = make_uint2(((uint32_t)b0 & 511u) | ((uint32_t)b1 & 511u) << 16, ((uint32_t)b2 & 511u) | ((uint32_t)b3 & 511u) << 16);
}

// grid (cdiv(P/8, 256), n_enc), unscaled profiles: eight pixels per thread. The interpolation of a byte is
// c1 + ((c2 - c1) >> 1) = floor((c1 + c2) / 2), one halving add on the packed word; the eight table look-ups are issued
// together and a single vote decides whether anybody missed (rare once the table is warm) - only then does the warp
// enter the cooperative evaluation. P is a multiple of 16 (width and height are multiples of 4).
// FULL: the table holds every colour (lut_fill_k): no miss check.
template <bool FULL>
__global__ void __launch_bounds__(256) quantize8_k(const SrcPair* __restrict__ src, uint32_t P, const uint32_t* __restrict__ pal, int npal,
                                                   uint16_t* __restrict__ lut, uint16_t* __restrict__ entries) {
    __shared__ uint32_t spal[FULL ? 1 : 512];
    if (!FULL) {
        for (int k = threadIdx.x; k < 512; k += blockDim.x) spal[k] = k < npal ? pal[k] : 0u;
        __syncthreads();
    }
    const uint32_t f = blockIdx.y;
    const SrcPair sp = src[f];
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;  // group of 8 pixels
    const bool valid = g * 8 < P;
    uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (valid) {
        const uint4* pa = reinterpret_cast<const uint4*>(sp.a) + 2 * (size_t)g;
        const uint4 a0 = __ldcs(pa), a1 = __ldcs(pa + 1);   // streamed once: evict-first, the colour table keeps the L2
        c[0] = a0.x; c[1] = a0.y; c[2] = a0.z; c[3] = a0.w; c[4] = a1.x; c[5] = a1.y; c[6] = a1.z; c[7] = a1.w;
        if (sp.b) {
            const uint4* pb = reinterpret_cast<const uint4*>(sp.b) + 2 * (size_t)g;
            const uint4 b0 = __ldcs(pb), b1 = __ldcs(pb + 1);
            c[0] = __vhaddu4(c[0], b0.x); c[1] = __vhaddu4(c[1], b0.y); c[2] = __vhaddu4(c[2], b0.z); c[3] = __vhaddu4(c[3], b0.w);
            c[4] = __vhaddu4(c[4], b1.x); c[5] = __vhaddu4(c[5], b1.y); c[6] = __vhaddu4(c[6], b1.z); c[7] = __vhaddu4(c[7], b1.w);
        }
    }
    uint32_t e[8];
    bool miss = false;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        c[k] &= 0xFFFFFFu;
        e[k] = valid ? lut[c[k]] : 0u;
    }
    if (!FULL) {
#pragma unroll
        for (int k = 0; k < 8; k++) miss |= e[k] == LUT_EMPTY;
        if (__any_sync(0xffffffffu, miss)) {
#pragma unroll
            for (int k = 0; k < 8; k++) e[k] = lut_entry(c[k], valid, lut, spal, npal);
        }
    }
    if (valid)
        __stcs(reinterpret_cast<uint4*>(entries + (size_t)f * P) + g, make_uint4(e[0] | e[1] << 16, e[2] | e[3] << 16, e[4] | e[5] << 16, e[6] | e[7] << 16));
}

// grid (cdiv(P/4, 256), n_enc). map (nullable): coded pixel -> source pixel (nearest-neighbour downscale, E13)
__global__ void __launch_bounds__(256) quantize_k(const SrcPair* __restrict__ src, const uint32_t* __restrict__ map, uint32_t P,
                                                  const uint32_t* __restrict__ pal, int npal, uint16_t* __restrict__ lut,
                                                  uint16_t* __restrict__ entries) {
    __shared__ uint32_t spal[512];
    for (int k = threadIdx.x; k < 512; k += blockDim.x) spal[k] = k < npal ? pal[k] : 0u;
    __syncthreads();
    const uint32_t f = blockIdx.y;
    const SrcPair sp = src[f];
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels
    const bool valid = g * 4 < P;
    uint32_t c[4] = {0, 0, 0, 0};
    if (valid) {
        if (map) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t s = map[g * 4 + k];
                uint32_t a = __ldg(sp.a + s);
                c[k] = sp.b ? interp_px(a, __ldg(sp.b + s)) : a;
            }
        } else {
            uint4 a = __ldg(reinterpret_cast<const uint4*>(sp.a) + g);
            c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w;
            if (sp.b) {
                uint4 b = __ldg(reinterpret_cast<const uint4*>(sp.b) + g);
                c[0] = interp_px(a.x, b.x); c[1] = interp_px(a.y, b.y);
                c[2] = interp_px(a.z, b.z); c[3] = interp_px(a.w, b.w);
            }
        }
    }
    uint32_t e[4];
#pragma unroll
    for (int k = 0; k < 4; k++) e[k] = lut_entry(c[k] & 0xFFFFFFu, valid, lut, spal, npal);
    if (valid) {
        uint2 o = make_uint2(e[0] | e[1] << 16, e[2] | e[3] << 16);
        reinterpret_cast<uint2*>(entries + (size_t)f * P)[g] = o;
    }
}

// ---------------------------------------------------------------------------
// AGMV_CompareFrameSimilarity (src/agmv_utils.c:920-947), used by AGMV_EncodeVideo to gate frame merging: number of
// pixels of two frames whose grey value (u8)((r+g+b)/3.0f) is equal. (r+g+b)/3.0f truncates to the integer quotient
// for every sum in 0..765 (a multiple of 3 divides exactly, anything else stays strictly between two integers), so the
// float expression is evaluated as an integer division here. grid (blocks, pairs).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) similarity_k(const SrcPair* __restrict__ pairs, const uint32_t* __restrict__ map, uint32_t P,
                                                    unsigned long long* __restrict__ counts) {
    const SrcPair sp = pairs[blockIdx.y];
    uint32_t acc = 0;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        const uint32_t s = map ? map[p] : p;
        const uint32_t a = __ldg(sp.a + s), b = __ldg(sp.b + s);
        const uint32_t ga = (((a >> 16) & 255u) + ((a >> 8) & 255u) + (a & 255u)) / 3u;
        const uint32_t gb = (((b >> 16) & 255u) + ((b >> 8) & 255u) + (b & 255u)) / 3u;
        acc += ga == gb;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane_id() == 0 && acc) atomicAdd(&counts[blockIdx.y], (unsigned long long)acc);
}

// ---------------------------------------------------------------------------
// K3: 4x4 block classification and bitstream assembly.
// AGMV_CompareIFrameBlock / AGMV_ComparePFrameBlock (src/agmv_encode.c:240-352)
// and AGMV_Assemble{I,P}FrameBitstream (:354-527). Record byte: type << 6 | len.
// ---------------------------------------------------------------------------

// |a - b| <= 2 on each of the three colour bytes (top byte 0 on both sides): per-byte absolute difference and compare on the
// packed word (classify_k was issue-bound at 73 % with the scalar form, profiles/r02_ncu_summary.txt)
__device__ __forceinline__ bool within2(uint32_t a, uint32_t b) { return __vcmpgtu4(__vabsdiffu4(a, b), 0x02020202u) == 0u; }
__device__ __forceinline__ uint32_t code_len(uint32_t e, int dual) { return (dual && (e & 255u) >= 127u) ? 2u : 1u; }

struct EntPair { const uint16_t* ent; const uint16_t* ient; };  // ient == nullptr: I-frame

__device__ __forceinline__ void load_block(const uint16_t* base, uint32_t W, uint32_t x, uint32_t y, uint32_t e[16]) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint2 v = *reinterpret_cast<const uint2*>(base + (size_t)(y + j) * W + x);
        e[j * 4 + 0] = v.x & 0xFFFFu; e[j * 4 + 1] = v.x >> 16;
        e[j * 4 + 2] = v.y & 0xFFFFu; e[j * 4 + 3] = v.y >> 16;
    }
}

// grid (cdiv(B,256), n_enc): writes one record byte per block
__global__ void __launch_bounds__(256) classify_k(const EntPair* __restrict__ fr, uint32_t W, uint32_t H, const uint32_t* __restrict__ pal,
                                                  int dual, uint8_t* __restrict__ rec) {
    __shared__ uint32_t spal[512];
    for (int k = threadIdx.x; k < 512; k += blockDim.x) spal[k] = pal[k];
    __syncthreads();
    const uint32_t bw = W >> 2, B = bw * (H >> 2);
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t f = blockIdx.y;
    const EntPair ep = fr[f];
    const uint32_t x = (b % bw) * 4, y = (b / bw) * 4;
    uint32_t e[16];
    load_block(ep.ent, W, x, y, e);
    const uint32_t c0 = spal[e[0]];
    int cfill = 0, ccopy = 0;
    uint32_t nlen = 1;
    if (ep.ient) {
        uint32_t ie[16];
        load_block(ep.ient, W, x, y, ie);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t bc = spal[e[k]];
            cfill += within2(c0, bc);
            ccopy += within2(bc, spal[ie[k]]);
            nlen += code_len(e[k], dual);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            cfill += within2(c0, spal[e[k]]);
            nlen += code_len(e[k], dual);
        }
    }
    uint32_t r;
    if (ep.ient && ccopy >= COPY_COUNT) r = BT_COPY << 6 | 1u;
    else if (cfill >= FILL_COUNT) r = BT_FILL << 6 | (1u + code_len(e[0], dual));
    else r = BT_NORMAL << 6 | nlen;
    rec[(size_t)f * B + b] = (uint8_t)r;
}

__device__ __forceinline__ uint32_t put_code(uint8_t* out, uint32_t pos, uint32_t e, int dual) {
    uint32_t idx = e & 255u, palbit = (e >> 8) & 1u;
    if (!dual) { out[pos++] = (uint8_t)idx; return pos; }
    if (idx < 127u) out[pos++] = (uint8_t)(palbit << 7 | idx);
    else { out[pos++] = (uint8_t)(palbit << 7 | 127u); out[pos++] = (uint8_t)idx; }
    return pos;
}

__global__ void __launch_bounds__(256) emit_k(const EntPair* __restrict__ fr, uint32_t W, uint32_t H, int dual, const uint8_t* __restrict__ rec,
                                              const uint32_t* __restrict__ boff, uint8_t* __restrict__ bs) {
    const uint32_t bw = W >> 2, B = bw * (H >> 2);
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t f = blockIdx.y;
    const uint32_t r = rec[(size_t)f * B + b];
    uint32_t pos = boff[(size_t)f * B + b];
    const uint32_t type = r >> 6;
    if (type == BT_COPY) { bs[pos] = COPY_FLAG; return; }
    const uint16_t* ent = fr[f].ent;
    const uint32_t x = (b % bw) * 4, y = (b / bw) * 4;
    if (type == BT_FILL) {
        bs[pos++] = FILL_FLAG;
        put_code(bs, pos, ent[(size_t)y * W + x], dual);
        return;
    }
    uint32_t e[16];
    load_block(ent, W, x, y, e);
    bs[pos++] = NORMAL_FLAG;
#pragma unroll
    for (int k = 0; k < 16; k++) pos = put_code(bs, pos, e[k], dual);
}

struct RecLen {
    const uint8_t* rec;
    __device__ uint32_t operator()(uint32_t i) const { return rec[i] & 63u; }
};

// frame starts: fs[f] = boff[f*B], fs[F] = grand total (left in scan_ws[ntiles] by device_scan)
__global__ void frame_starts_k(const uint32_t* __restrict__ boff, uint32_t B, uint32_t F, const uint32_t* __restrict__ total,
                               uint32_t* __restrict__ fs) {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < F) fs[f] = boff[(size_t)f * B];
    else if (f == F) fs[F] = *total;
}

}  // namespace agmvb
