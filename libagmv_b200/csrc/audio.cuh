// Audio chunk codec: AGMV_CompressAudio (src/agmv_encode.c:659-705) and the sample loop of
// AGMV_DecodeAudioChunk (src/agmv_decode.c:412-453) as streaming kernels. Both are pure per-sample maps, HBM-bound:
// 2 B in + 1 B out per sample (compress), 1 B in + 2 B out (expand); 16-byte accesses, grid-stride.
#pragma once
#include "common.cuh"

namespace agmvb {

// One 16-bit sample -> one byte. The reference tries three codes and keeps the closest:
//   ssqrt1 = (u8)sqrt(s) rounded up to even, ssqrt2 = (u8)round(sqrt(s)) rounded up to even, shift = (s >> 8) rounded up to odd;
// the decoder squares even bytes and shifts odd bytes left by 8. All of it is exact in integers:
//   (u8)sqrt(s) = isqrt(s); round(sqrt(s)) = isqrt(s) + (s > k*k + k) (sqrt(s) >= k + 1/2 <=> s >= k*k + k + 1/4);
//   round = 256 for s > 65280, stored in a u8 as 0 (what the x86 build of the reference does; pinned by the known answers
//   of all 65 536 inputs, tests/golden/audio_compress16.bin); roundUpEven wraps 255 -> 0 in its u8 (:15-21).
// Selection (:683-696): dist = min(dist1, dist3) - dist2 only competes on a tie with dist3.
__device__ __forceinline__ uint32_t audio_compress_sample(uint32_t s) {
    uint32_t k = (uint32_t)sqrtf((float)s);  // s < 2^16: exact to within one, fixed below
    k -= (k * k > s);
    k += ((k + 1) * (k + 1) <= s);
    const uint32_t r = (k + (s > k * k + k)) & 255u;
    const uint32_t s1 = (k + 1) & 0xFEu, s2 = (r + 1) & 0xFEu, sh = (s >> 8) | 1u;
    const int d1 = abs((int)(s1 * s1) - (int)s), d2 = abs((int)(s2 * s2) - (int)s), d3 = abs((int)(sh << 8) - (int)s);
    const int d = min(d1, d3);
    return d == d1 ? s1 : (d == d2 ? s2 : sh);
}

// even byte -> AGMV_SQR_TABLE[b] = b*b, odd byte -> AGMV_SHIFT_TABLE[b] = b << 8 (src/agmv_decode.c:21-89, :433-438)
__device__ __forceinline__ uint32_t audio_expand_sample(uint32_t b) { return (b & 1u) ? (b << 8) : b * b; }

// pcm (u16, 16-byte aligned) -> atsample; 16 samples per thread step
__global__ void __launch_bounds__(256) audio_compress16_k(const uint16_t* __restrict__ pcm, uint64_t n, uint8_t* __restrict__ at) {
    const uint64_t n16 = n >> 4;
    const uint4* in = reinterpret_cast<const uint4*>(pcm);
    uint4* out = reinterpret_cast<uint4*>(at);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            o[q] = audio_compress_sample(w[2 * q] & 0xFFFFu) | audio_compress_sample(w[2 * q] >> 16) << 8 |
                   audio_compress_sample(w[2 * q + 1] & 0xFFFFu) << 16 | audio_compress_sample(w[2 * q + 1] >> 16) << 24;
        }
        out[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (blockIdx.x == 0)
        for (uint64_t i = (n16 << 4) + threadIdx.x; i < n; i += blockDim.x) at[i] = (uint8_t)audio_compress_sample(pcm[i]);
}

// The same map through a table: all 65 536 answers (audio_lut_k, once per context, 64 KB) staged into shared memory by every
// CTA of a persistent grid, then one shared-memory byte load per sample instead of ~45 ALU / SFU instructions - the ALU pipe
// is what bounds audio_compress16_k (ncu: 71 % ALU, 28 % of HBM peak). Used for tracks long enough to amortise the staging.
__global__ void __launch_bounds__(256) audio_lut_k(uint8_t* __restrict__ lut) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < 65536u) lut[s] = (uint8_t)audio_compress_sample(s);
}

constexpr int AUDIO_LUT_THREADS = 512;
__global__ void __launch_bounds__(AUDIO_LUT_THREADS) audio_compress16_lut_k(const uint16_t* __restrict__ pcm, uint64_t n, const uint8_t* __restrict__ lut,
                                                                           uint8_t* __restrict__ at) {
    extern __shared__ uint4 lut_s4[];
    const uint8_t* t = reinterpret_cast<const uint8_t*>(lut_s4);
    for (int i = threadIdx.x; i < 4096; i += AUDIO_LUT_THREADS) lut_s4[i] = __ldg(reinterpret_cast<const uint4*>(lut) + i);
    __syncthreads();
    const uint64_t n16 = n >> 4, stride = (uint64_t)gridDim.x * blockDim.x;
    const uint4* in = reinterpret_cast<const uint4*>(pcm);
    uint4* out = reinterpret_cast<uint4*>(at);
    auto code4 = [&](uint32_t w0, uint32_t w1) { return (uint32_t)t[w0 & 0xFFFFu] | (uint32_t)t[w0 >> 16] << 8 | (uint32_t)t[w1 & 0xFFFFu] << 16 | (uint32_t)t[w1 >> 16] << 24; };
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; i + stride < n16; i += 2 * stride) {  // two independent 32-byte reads in flight per thread
        const uint4 a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1), c = __ldg(in + 2 * (i + stride)), d = __ldg(in + 2 * (i + stride) + 1);
        out[i] = make_uint4(code4(a.x, a.y), code4(a.z, a.w), code4(b.x, b.y), code4(b.z, b.w));
        out[i + stride] = make_uint4(code4(c.x, c.y), code4(c.z, c.w), code4(d.x, d.y), code4(d.z, d.w));
    }
    if (i < n16) {
        const uint4 a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1);
        out[i] = make_uint4(code4(a.x, a.y), code4(a.z, a.w), code4(b.x, b.y), code4(b.z, b.w));
    }
    if (blockIdx.x == 0)
        for (uint64_t k = (n16 << 4) + threadIdx.x; k < n; k += blockDim.x) at[k] = t[pcm[k]];
}

// atsample -> pcm (u16)
__global__ void __launch_bounds__(256) audio_expand16_k(const uint8_t* __restrict__ at, uint64_t n, uint16_t* __restrict__ pcm) {
    const uint64_t n16 = n >> 4;
    const uint4* in = reinterpret_cast<const uint4*>(at);
    uint4* out = reinterpret_cast<uint4*>(pcm);
    auto put = [&](uint64_t i, const uint4 a) {
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            o[2 * q] = audio_expand_sample(w[q] & 255u) | audio_expand_sample((w[q] >> 8) & 255u) << 16;
            o[2 * q + 1] = audio_expand_sample((w[q] >> 16) & 255u) | audio_expand_sample(w[q] >> 24) << 16;
        }
        __stcs(out + 2 * i, make_uint4(o[0], o[1], o[2], o[3]));
        __stcs(out + 2 * i + 1, make_uint4(o[4], o[5], o[6], o[7]));
    };
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {  // four independent reads in flight per thread
        const uint4 a = __ldg(in + i), b = __ldg(in + i + stride), c = __ldg(in + i + 2 * stride), d = __ldg(in + i + 3 * stride);
        put(i, a); put(i + stride, b); put(i + 2 * stride, c); put(i + 3 * stride, d);
    }
    for (; i < n16; i += stride) put(i, __ldg(in + i));
    if (blockIdx.x == 0)
        for (uint64_t i = (n16 << 4) + threadIdx.x; i < n; i += blockDim.x) pcm[i] = (uint16_t)audio_expand_sample(at[i]);
}

// The audio half of a stream's decode loop (src/agmv_decode.c:583-587): chunk c holds len[c] sample bytes at file offset
// off[c]; its samples land at dst[c] (= the sum of the earlier chunks' sizes, audio_track->start_point). Bytes past the end
// of the file read as 0xFF (fgetc's EOF stored in a u8). bits == 16: expand to u16, else copy the byte (pcm8).
__global__ void __launch_bounds__(256) audio_track_k(const uint8_t* __restrict__ file, uint64_t file_len, const uint64_t* __restrict__ off,
                                                     const uint32_t* __restrict__ len, const uint64_t* __restrict__ dst, int bits,
                                                     uint16_t* __restrict__ pcm16, uint8_t* __restrict__ pcm8) {
    const uint32_t c = blockIdx.y;
    const uint64_t o = off[c], d = dst[c];
    const uint32_t n = len[c];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t b = o + i < file_len ? file[o + i] : 0xFFu;
        if (bits == 16) pcm16[d + i] = (uint16_t)audio_expand_sample(b);
        else pcm8[d + i] = (uint8_t)b;
    }
}

}  // namespace agmvb
