// One stable 8-bit counting-sort pass over functor-defined elements.
//
//   radix_hist_k     per-tile digit counts           -> tile_hist[digit * ntiles + tile]
//   (device_scan)    exclusive sum in digit-major order gives every (digit, tile) its base
//   radix_scatter_k  stable rank inside the tile + base -> mv(src, dst)
//
// Stability inside a tile comes from the work layout: each warp owns 512
// consecutive elements and sweeps them in 16 rounds of 32; within a round
// __match_any_sync ranks equal digits by lane, per-warp digit counters carry
// the rank from round to round, and a prefix over the 8 warps orders the warps.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace agmvb {

constexpr int RX_THREADS = 256;
constexpr int RX_WARPS = RX_THREADS / 32;
constexpr int RX_ROUNDS = 16;
constexpr int RX_WARP_SPAN = 32 * RX_ROUNDS;
constexpr int RX_TILE = RX_THREADS * RX_ROUNDS;  // 4096

template <class DigitF>
__global__ void __launch_bounds__(RX_THREADS) radix_hist_k(DigitF dg, uint32_t n, uint32_t ntiles, uint32_t* __restrict__ tile_hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * RX_TILE + warp * RX_WARP_SPAN + lane_id();
#pragma unroll 4
    for (int r = 0; r < RX_ROUNDS; r++) {
        uint32_t i = base + r * 32;
        bool valid = i < n;
        uint32_t key = valid ? dg(i) : 256u + lane_id();
        unsigned peers = __match_any_sync(0xffffffffu, key);
        if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&h[key], __popc(peers));
    }
    __syncthreads();
    tile_hist[threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

template <class DigitF, class MoveF>
__global__ void __launch_bounds__(RX_THREADS) radix_scatter_k(DigitF dg, MoveF mv, uint32_t n, uint32_t ntiles,
                                                              const uint32_t* __restrict__ tile_off) {
    __shared__ uint32_t wc[RX_WARPS][256];
    __shared__ uint32_t goff[256];
#pragma unroll
    for (int w = 0; w < RX_WARPS; w++) wc[w][threadIdx.x] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * RX_TILE + warp * RX_WARP_SPAN + lane_id();
    uint32_t packed[RX_ROUNDS];  // digit << 16 | rank inside the warp's span
#pragma unroll
    for (int r = 0; r < RX_ROUNDS; r++) {
        uint32_t i = base + r * 32;
        bool valid = i < n;
        uint32_t key = valid ? dg(i) : 256u + lane_id();
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
    {
        uint32_t run = 0;
        const int d = threadIdx.x;
#pragma unroll
        for (int w = 0; w < RX_WARPS; w++) {
            uint32_t t = wc[w][d];
            wc[w][d] = run;
            run += t;
        }
        goff[d] = tile_off[d * ntiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RX_ROUNDS; r++) {
        uint32_t i = base + r * 32;
        if (i < n) {
            uint32_t d = packed[r] >> 16, rk = packed[r] & 0xffffu;
            mv(i, goff[d] + wc[warp][d] + rk);
        }
    }
}

struct LoadU32 {
    const uint32_t* p;
    __device__ uint32_t operator()(uint32_t i) const { return p[i]; }
};
struct StoreU32 {
    uint32_t* p;
    __device__ void operator()(uint32_t i, uint32_t v) const { p[i] = v; }
};

// tile_hist: 256 * ntiles words; scan_ws: cdiv(256*ntiles, SCAN_TILE) + 1 words
template <class DigitF, class MoveF>
inline void radix_pass(DigitF dg, MoveF mv, uint32_t n, uint32_t* tile_hist, uint32_t* scan_ws, LaunchCtx& lc) {
    cudaStream_t st = lc.st;
    if (n == 0) return;
    uint32_t nt = (n + RX_TILE - 1) / RX_TILE;
    KL(lc, KC_RX_HIST, (radix_hist_k<DigitF><<<nt, RX_THREADS, 0, st>>>(dg, n, nt, tile_hist)));
    device_scan<SumOp, true>(LoadU32{tile_hist}, StoreU32{tile_hist}, 256u * nt, scan_ws, lc, KC_RX_SCAN);
    KL(lc, KC_RX_SCATTER, (radix_scatter_k<DigitF, MoveF><<<nt, RX_THREADS, 0, st>>>(dg, mv, n, nt, tile_hist)));
}

}  // namespace agmvb
