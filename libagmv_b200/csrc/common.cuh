// Shared definitions for the sm_100a kernels behind include/agmv_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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
            (void)cudaGetLastError(); /* reported through the return code: do not leave it for the next CUDA user */ \
            return ERR_CUDA;                                                                 \
        }                                                                                    \
    } while (0)

// ---- launch bookkeeping -------------------------------------------------------
// Every kernel launch goes through LaunchCtx: it counts launches (bench.py's
// gpu_launches) and, when profiling is on, brackets the launch with CUDA events
// on the launching stream so per-kernel-class device time can be read back live.
enum KClass {
    KC_HIST = 0, KC_PALETTE, KC_QUANT, KC_CLASSIFY, KC_BLOCKSCAN, KC_EMIT, KC_LZ_INIT, KC_RX_HIST, KC_RX_SCAN, KC_RX_SCATTER,
    KC_LZ_LINK, KC_LZ_LINK3, KC_LZ_LEVEL, KC_LZ_PARSE, KC_LZ_PACK, KC_LZ_CHUNK, KC_EXPAND, KC_STALE, KC_INDEX, KC_RECON, KC_CHECKSUM, KC_MISC,
    KC_LZ77, KC_AUDIO, KC_COUNT
};
inline const char* kclass_name(int c) {
    static const char* n[] = {"hist", "palette", "quantize", "classify", "block_scan", "emit", "lz_init", "rx_hist", "rx_scan",
                              "rx_scatter", "lz_link", "lz_link3", "lz_level", "lz_parse", "lz_pack", "lz_chunk", "expand", "stale", "index",
                              "reconstruct", "checksum", "misc", "lz77", "audio"};
    return c >= 0 && c < KC_COUNT ? n[c] : "?";
}
struct LaunchCtx {
    cudaStream_t st = nullptr;
    uint64_t launches = 0;
    bool prof = false;
    struct Rec { int cls; cudaEvent_t a, b; };
    Rec* recs = nullptr;
    size_t nrec = 0, cap = 0;
    cudaEvent_t* pool = nullptr;
    size_t npool = 0, pool_cap = 0;
    cudaEvent_t get_event() {
        if (npool == pool_cap) {
            size_t nc = pool_cap ? pool_cap * 2 : 4096;
            pool = (cudaEvent_t*)realloc(pool, nc * sizeof(cudaEvent_t));
            for (size_t k = pool_cap; k < nc; k++) cudaEventCreate(&pool[k]);
            pool_cap = nc;
        }
        return pool[npool++];
    }
    void begin(int cls) {
        if (!prof) return;
        if (nrec == cap) { cap = cap ? cap * 2 : 4096; recs = (Rec*)realloc(recs, cap * sizeof(Rec)); }
        recs[nrec].cls = cls;
        recs[nrec].a = get_event();
        recs[nrec].b = get_event();
        cudaEventRecord(recs[nrec].a, st);
    }
    void end(int) {
        launches++;
        if (!prof) return;
        cudaEventRecord(recs[nrec].b, st);
        nrec++;
    }
};
#define KL(lc, cls, ...) do { (lc).begin(cls); __VA_ARGS__; (lc).end(cls); } while (0)

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace agmvb
