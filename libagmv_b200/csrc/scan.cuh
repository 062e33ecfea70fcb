// Device-wide scans over functor inputs (three launches: tile reduce, scan of
// the tile partials by one CTA, tile apply). Used for block-record offsets
// (sum), radix bucket offsets (sum) and group starts (running max).
#pragma once
#include "common.cuh"

namespace agmvb {

struct SumOp {
    __device__ static uint32_t id() { return 0u; }
    __device__ static uint32_t op(uint32_t a, uint32_t b) { return a + b; }
};
struct MaxOp {
    __device__ static uint32_t id() { return 0u; }
    __device__ static uint32_t op(uint32_t a, uint32_t b) { return a > b ? a : b; }
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ROUNDS = 16;                       // items per thread
constexpr int SCAN_WARP_SPAN = 32 * SCAN_ROUNDS;      // 512 consecutive items per warp
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ROUNDS; // 4096

template <class Op>
__device__ __forceinline__ uint32_t warp_inclusive(uint32_t x) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane_id() >= (unsigned)d) x = Op::op(y, x);
    }
    return x;
}

template <class Op, class InF>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_k(InF in, uint32_t n, uint32_t* __restrict__ partial) {
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    uint32_t base = blockIdx.x * SCAN_TILE;
    uint32_t acc = Op::id();
#pragma unroll 4
    for (int k = 0; k < SCAN_ROUNDS; k++) {
        uint32_t i = base + k * SCAN_THREADS + threadIdx.x;
        if (i < n) acc = Op::op(acc, in(i));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = Op::op(acc, __shfl_xor_sync(0xffffffffu, acc, d));
    if (lane_id() == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = Op::id();
        for (int w = 0; w < SCAN_THREADS / 32; w++) t = Op::op(t, wsum[w]);
        partial[blockIdx.x] = t;
    }
}

// exclusive scan of partial[0..np) in place; partial[np] receives the total
template <class Op>
__global__ void __launch_bounds__(1024) scan_partials_k(uint32_t* __restrict__ partial, uint32_t np) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = Op::id();
    __syncthreads();
    for (uint32_t base = 0; base < np; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < np ? partial[i] : Op::id();
        uint32_t inc = warp_inclusive<Op>(x);
        if (lane_id() == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t t = warp_inclusive<Op>(wsum[threadIdx.x]);
            wsum[threadIdx.x] = t;
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t wpre = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : Op::id();
        uint32_t incl = Op::op(carry, Op::op(wpre, inc));
        // exclusive value = everything before me
        uint32_t prev = __shfl_up_sync(0xffffffffu, inc, 1);
        uint32_t excl = Op::op(carry, lane_id() == 0 ? wpre : Op::op(wpre, prev));
        if (i < np) partial[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[np] = carry_s;
}

// Inclusive (or exclusive, sum only) scan applied per tile with the tile's
// carry-in. Each warp owns 512 consecutive items and sweeps them in 16
// coalesced rounds, so the order of items is preserved.
template <class Op, bool EXCLUSIVE, class InF, class OutF>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_k(InF in, OutF out, uint32_t n, const uint32_t* __restrict__ partial) {
    __shared__ uint32_t wtot[SCAN_THREADS / 32];
    const int warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * SCAN_TILE + warp * SCAN_WARP_SPAN + lane_id();
    uint32_t v[SCAN_ROUNDS], x0[SCAN_ROUNDS];
    uint32_t carry = Op::id();
#pragma unroll
    for (int r = 0; r < SCAN_ROUNDS; r++) {
        uint32_t i = base + r * 32;
        uint32_t x = i < n ? in(i) : Op::id();
        x0[r] = x;
        x = warp_inclusive<Op>(x);
        x = Op::op(carry, x);
        v[r] = x;
        carry = __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane_id() == 0) wtot[warp] = carry;
    __syncthreads();
    uint32_t pre = partial[blockIdx.x];
    for (int w = 0; w < warp; w++) pre = Op::op(pre, wtot[w]);
#pragma unroll
    for (int r = 0; r < SCAN_ROUNDS; r++) {
        uint32_t i = base + r * 32;
        if (i < n) {
            uint32_t incl = Op::op(pre, v[r]);
            out(i, EXCLUSIVE ? incl - x0[r] : incl);
        }
    }
}

// ws must hold cdiv(n, SCAN_TILE) + 1 words. After the call ws[ntiles] holds
// the grand total (sum) / maximum.
template <class Op, bool EXCLUSIVE, class InF, class OutF>
inline void device_scan(InF in, OutF out, uint32_t n, uint32_t* ws, LaunchCtx& lc, int cls) {
    cudaStream_t st = lc.st;
    if (n == 0) {
        cudaMemsetAsync(ws, 0, sizeof(uint32_t), st);
        return;
    }
    uint32_t nt = (n + SCAN_TILE - 1) / SCAN_TILE;
    KL(lc, cls, (scan_reduce_k<Op, InF><<<nt, SCAN_THREADS, 0, st>>>(in, n, ws)));
    KL(lc, cls, (scan_partials_k<Op><<<1, 1024, 0, st>>>(ws, nt)));
    KL(lc, cls, (scan_apply_k<Op, EXCLUSIVE, InF, OutF><<<nt, SCAN_THREADS, 0, st>>>(in, out, n, ws)));
}

}  // namespace agmvb
