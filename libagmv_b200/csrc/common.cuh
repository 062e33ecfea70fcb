// Shared definitions for the sm_100a kernels behind include/agmv_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace agmvb {

// reference constants (include/agmv_defines.h:49-53)
constexpr uint8_t FILL_FLAG = 0x4E;
constexpr uint8_t NORMAL_FLAG = 0x2F;
constexpr uint8_t COPY_FLAG = 0x5E;
constexpr int FILL_COUNT = 14;
constexpr int COPY_COUNT = 13;
// src/agmv_encode.c:101-104
constexpr int LZ_WINDOW = 65535;
constexpr int LZ_MAXLEN = 15;
constexpr int LZ_MINLEN = 3;

// block record types shared by the encoder (K3) and the decoder (D3)
constexpr uint32_t BT_FILL = 1, BT_COPY = 2, BT_NORMAL = 3;

constexpr uint32_t EMPTY32 = 0xFFFFFFFFu;
constexpr uint16_t LUT_EMPTY = 0xFFFFu;

enum { OPT_I = 1, OPT_II, OPT_III, OPT_ANIM, OPT_GBA_I, OPT_GBA_II, OPT_GBA_III, OPT_NDS };
enum { Q_HIGH = 1, Q_MID, Q_LOW };
enum { COMP_LZSS = 1, COMP_LZ77 };

// error codes of the C-ABI (0 = ok); the reference's own decode codes 1..3 are kept
// (include/agmv_defines.h:37-42) and CUDA failures map to 3 as SURVEY.md 8b suggests.
enum { OK = 0, ERR_HEADER = 1, ERR_FILE = 2, ERR_MEMORY = 3, ERR_ARG = 4, ERR_CUDA = 5, ERR_UNSUPPORTED = 6 };

__host__ __device__ inline uint32_t max_clr(int q) { return q == Q_MID ? 131071u : (q == Q_LOW ? 65535u : 524287u); }
__host__ __device__ inline bool opt_is_dual(int o) { return !(o == OPT_II || o == OPT_ANIM || o == OPT_GBA_II); }
__host__ __device__ inline bool opt_is_light(int o) { return !(o == OPT_I || o == OPT_ANIM || o == OPT_GBA_I || o == OPT_GBA_II); }

// src/agmv_utils.c:695-742
__host__ __device__ inline uint32_t quantize_color(uint32_t c, int q) {
    uint32_t r = (c >> 16) & 255, g = (c >> 8) & 255, b = c & 255;
    if (q == Q_MID) return (r >> 3) << 12 | (g >> 2) << 6 | (b >> 2);
    if (q == Q_LOW) return (r >> 3) << 11 | (g >> 2) << 5 | (b >> 3);
    return (r >> 2) << 13 | (g >> 2) << 7 | (b >> 1);
}

inline int cdiv(size_t a, size_t b) { return (int)((a + b - 1) / b); }

#define AGMVB_CUDA_OK(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            snprintf(ctx->err, sizeof ctx->err, "%s:%d %s: %s", __FILE__, __LINE__, #expr,   \
                     cudaGetErrorString(_e));                                                \
            return ERR_CUDA;                                                                 \
        }                                                                                    \
    } while (0)

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace agmvb
