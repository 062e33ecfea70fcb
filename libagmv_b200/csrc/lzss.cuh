// K4: exact LZSS for a batch of frame bitstreams (reference: AGMV_LZSS,
// src/agmv_encode.c:106-177, bit writer src/agmv_utils.c:86-112, chunk framing
// src/agmv_encode.c:549-624).
//
// The reference does a brute-force search per parse position: longest match
// (3..15 bytes, overlap allowed) inside the previous 65535 bytes and, among
// equal lengths, the EARLIEST start. That is O(n * 65535) per frame. Here the
// same answer comes from occurrence chains, one level per match length
// (lzchain_core.h states the algorithm, lzchain.cuh holds its kernels): no
// position is sorted or moved, a level is one streaming pass plus a few gathers.
//
// The greedy parse itself is a pointer chase (i += len or 1). It is resolved
// tile-parallel (orbit.cuh): every 1024-position tile is walked speculatively
// from each of the 15 possible entry offsets, one warp per frame then chains the
// tiles, and a last pass re-walks each tile from its real entry writing the
// running bit cursor. Bits are packed LSB-first with atomicOr on 32-bit words;
// csize uses the same float expression as the reference; the 8x0xFF trailer is
// written at data+csize so it clobbers the last partial byte exactly like the
// reference.
#pragma once
#include "common.cuh"
#include "orbit.cuh"
#include "scan.cuh"
#include <vector>
#include "lzchain.cuh"

namespace agmvb {

struct LzWork {
    uint32_t cap_n = 0, cap_frames = 0;
    uint8_t* bestlen = nullptr;          // per position: longest match length (0 = none), input of the parse
    uint32_t* lw[2] = {};                // link words of the current / next level, ping-pong (lw[1] doubles as the hash-level links)
    uint16_t* rsd = nullptr;             // distance of every position to the start of its byte run
    LzcItem* items = nullptr;            // (frame, range) work items of lzc_hashlink_k
    uint32_t n_items = 0;
    int hash_bits = LZC_HB;              // size of lzc_hashlink_k's table
    uint32_t* cframe = nullptr;          // frame of the first position of every LZC_WCHUNK-position chunk
    uint32_t* counters = nullptr;        // chunk counters of the persistent kernels (one per launch of a batch)
    uint32_t link3_blocks = 148 * 5, level_blocks = 148 * 6;   // resident blocks of the persistent walking kernels
    uint32_t* list15 = nullptr;          // per parse tile: its 15-byte matches (position in tile | bit cursor in tile << 10)
    uint8_t* cnt15 = nullptr;            // per parse tile: how many
    OrbitTables orb;                     // greedy-parse tables (cap_n / ORB_TILE + cap_frames tiles)
    OrbitSeg* segs = nullptr;            // cap_frames
    uint32_t* seg_len = nullptr;         // cap_frames
    uint32_t* out_words = nullptr;
    size_t out_words_cap = 0;
    uint32_t* wbase = nullptr;           // cap_frames + 1
    uint32_t* outbits = nullptr;         // cap_frames
    uint32_t* csize = nullptr;           // cap_frames
    uint32_t* chunk_off = nullptr;       // cap_frames + 1
    uint32_t stub_bytes = 8;             // empty 'AGAC' chunk after every frame chunk (AGMV_EncodeAGMV); 0 for the other encoders
    uint32_t audio_chunk = 0;            // bytes of audio per frame (stub_bytes = 8 + audio_chunk while a track is interleaved)
    const uint8_t* audio = nullptr;      // companded track on the device (AGMV_CompressAudio's atsample)
    uint64_t audio_size = 0;
};

// ---- greedy parse: orbit over bestlen[] (orbit.cuh) -------------------------
struct LzStep {
    __device__ static uint32_t step(uint32_t c) { return c >= (uint32_t)LZ_MINLEN ? c : 1u; }
    __device__ static uint32_t weight(uint32_t c) { return c >= (uint32_t)LZ_MINLEN ? 21u : 9u; }  // bits per token
};
// ---- token emission from the parse itself (bit writer: src/agmv_utils.c:86-112; token layout src/agmv_encode.c:146-165) ----
// The last step of the orbit re-walks every tile from its true entry with the running bit cursor in a register: the lane
// that walks a tile emits its literals and its matches shorter than 15 bytes on the spot (the level-15 word of a finished
// position holds its offset) and notes the tile's 15-byte matches - whose earliest start is the end of a chain walk - in a
// small per-tile list for lz_pack15_k. (Round 2 first wrote a bit cursor per visited position for a separate pack kernel:
// 141 M scattered 4-byte stores per step, an ECC sector fill each, then a dense read of that array.)
constexpr int LZ15_SLOTS = 72;   // >= ceil(1024 / 15) matches of 15 bytes start in a tile
__global__ void __launch_bounds__(ORB_LANES) lz_emit_mark_k(const uint8_t* __restrict__ code, const OrbitSeg* __restrict__ segs, uint32_t nseg,
                                                            const uint32_t* __restrict__ seg_len, uint32_t ntile, OrbitTables tb,
                                                            const uint8_t* __restrict__ bs, const uint32_t* __restrict__ lw15,
                                                            const uint32_t* __restrict__ wbase, uint32_t* __restrict__ out_words,
                                                            uint32_t* __restrict__ list15, uint8_t* __restrict__ cnt15) {
    const uint32_t tile = blockIdx.x * ORB_LANES + threadIdx.x;
    if (tile >= ntile) return;
    const uint32_t sg = orbit_seg_of(segs, nseg, tile);
    const OrbitSeg seg = segs[sg];
    const uint32_t len = seg_len[sg];
    const uint32_t t0 = (tile - seg.tile_base) * ORB_TILE;
    uint32_t n15 = 0;
    if (t0 < len) {
        const uint32_t end = min(t0 + ORB_TILE, len);
        const uint32_t wb = wbase[sg];
        const uint8_t* const d = bs + seg.off;
        const uint32_t* const lw = lw15 + seg.off;
        OrbitBytes c;
        c.init(code + seg.off);
        const uint32_t cum0 = tb.cumbase[tile];
        uint32_t i = t0 + tb.entry_tab[tile], cum = cum0;
        while (i < end) {
            const uint32_t l = c.at(i);
            if (l < (uint32_t)LZ_MINLEN) {
                lzc_emit(out_words, wb, cum, 1u | ((uint32_t)d[i] << 1), 9);
                cum += 9; i += 1;
            } else {
                if (l < (uint32_t)LZ_MAXLEN) lzc_emit(out_words, wb, cum, ((lw[i] & 0xFFFFu) << 1) | (l << 17), 21);
                else list15[(size_t)tile * LZ15_SLOTS + n15++] = (i - t0) | (cum - cum0) << 10;
                cum += 21; i += l;
            }
        }
    }
    cnt15[tile] = (uint8_t)n15;
}

// the 15-byte matches of the parse: each walks its level-15 chain to the end (lane refill as in the level kernels); a warp
// takes LZ15_GROUP consecutive tiles
constexpr int LZ15_GROUP = 8;
__global__ void __launch_bounds__(LZC_THREADS) lz_pack15_k(const OrbitSeg* __restrict__ segs, uint32_t nseg, uint32_t ntile, OrbitTables tb,
                                                           const uint32_t* __restrict__ lw15, const uint16_t* __restrict__ rsd,
                                                           const uint32_t* __restrict__ wbase, const uint32_t* __restrict__ list15,
                                                           const uint8_t* __restrict__ cnt15, uint32_t* __restrict__ out_words) {
    __shared__ uint32_t q[LZC_WARPS][LZ15_GROUP * LZ15_SLOTS];   // tile in group << 28 | list entry
    __shared__ uint32_t tpos[LZC_WARPS][LZ15_GROUP], tcum[LZC_WARPS][LZ15_GROUP], twb[LZC_WARPS][LZ15_GROUP];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t tile0 = (blockIdx.x * LZC_WARPS + warp) * LZ15_GROUP;
    if (tile0 >= ntile) return;
    uint32_t qn = 0;
    for (uint32_t g = 0; g < (uint32_t)LZ15_GROUP && tile0 + g < ntile; g++) {
        const uint32_t tile = tile0 + g, cnt = cnt15[tile];
        if (lane == 0) {
            const uint32_t sg = orbit_seg_of(segs, nseg, tile);
            tpos[warp][g] = (uint32_t)segs[sg].off + (tile - segs[sg].tile_base) * ORB_TILE;
            tcum[warp][g] = tb.cumbase[tile];
            twb[warp][g] = wbase[sg];
        }
        for (uint32_t k = lane; k < cnt; k += 32) q[warp][qn + k] = g << 28 | list15[(size_t)tile * LZ15_SLOTS + k];
        qn += cnt;
    }
    __syncwarp();
    uint32_t qi = 0, rel = 0, wb = 0;
    bool busy = false;
    LzcEndWalk wlk;
    for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, !busy);
        if (qi < qn && idle) {
            const uint32_t my = qi + __popc(idle & lanemask_lt());
            if (!busy && my < qn) {
                const uint32_t e = q[warp][my], g = e >> 28;
                const uint32_t p = tpos[warp][g] + (e & 1023u);
                rel = tcum[warp][g] + ((e >> 10) & 0x3FFFFu);
                wb = twb[warp][g];
                wlk.start(p, lw15[p]);   // a 15-byte match: the link is never 0
                busy = true;
            }
            qi += __popc(idle);
        }
        if (!__any_sync(0xffffffffu, busy)) break;
        if (busy && wlk.hop(lw15, rsd) != LZC_GO) {
            lzc_emit(out_words, wb, rel, (wlk.last << 1) | ((uint32_t)LZ_MAXLEN << 17), 21);
            busy = false;
        }
    }
}

__global__ void __launch_bounds__(1024) lz_finalize_k(uint32_t F, uint32_t stub, const uint32_t* __restrict__ total_bits,
                                                      uint32_t* __restrict__ outbits, uint32_t* __restrict__ csize,
                                                      uint32_t* __restrict__ chunk_off) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < F; base += 1024) {
        uint32_t f = base + threadIdx.x;
        uint32_t len = 0;
        if (f < F) {
            uint32_t ob = total_bits[f];
            uint32_t cs = (uint32_t)((float)(int)ob / 8.0f);  // src/agmv_encode.c:176
            outbits[f] = ob;
            csize[f] = cs;
            len = 24u + stub + cs;  // AGFC hdr 16 + payload + 8 x 0xFF (+ empty AGAC chunk 8)
        }
        uint32_t inc = warp_inclusive<SumOp>(len);
        if (lane_id() == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) wsum[threadIdx.x] = warp_inclusive<SumOp>(wsum[threadIdx.x]);
        __syncthreads();
        uint32_t pre = carry_s + ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0u);
        if (f < F) chunk_off[f] = pre + inc - len;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = pre + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_off[F] = carry_s;
}

// src/agmv_encode.c:549-551,567-585,622-624 and :707-717 (empty audio chunk)
__global__ void __launch_bounds__(256) lz_write_chunks_k(const uint32_t* __restrict__ fs, const uint32_t* __restrict__ csize,
                                                         const uint32_t* __restrict__ chunk_off, const uint32_t* __restrict__ wbase,
                                                         const uint32_t* __restrict__ out_words, uint32_t first_frame_count,
                                                         uint32_t stub, uint8_t* __restrict__ image, uint32_t audio_chunk = 0,
                                                         const uint8_t* __restrict__ audio = nullptr, uint64_t audio_size = 0) {
    const uint32_t f = blockIdx.y;
    const uint32_t cs = csize[f], len = 24u + stub + cs, usize = fs[f + 1] - fs[f];
    const uint8_t* pay = reinterpret_cast<const uint8_t*>(out_words + wbase[f]);
    uint8_t* dst = image + chunk_off[f];
    const uint32_t num = first_frame_count + f + 1;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < len; b += gridDim.x * blockDim.x) {
        uint8_t v;
        if (b < 16) {
            const uint32_t word = b < 4 ? 0x43464741u /* "AGFC" */ : (b < 8 ? num : (b < 12 ? usize : cs));
            v = (uint8_t)(word >> ((b & 3) * 8));
        } else if (b < 16 + cs) v = pay[b - 16];
        else if (b < 24 + cs) v = 0xFF;
        else {
            // AGMV_EncodeAudioChunk (:707-717): 'AGAC', size, then size bytes from atsample[start_point++]; frame number g of the
            // sequence starts at g * size. Reads past the track (the reference walks off its malloc) are defined as 0.
            const uint32_t k = b - 24 - cs;
            if (k < 4) v = (uint8_t)(0x43414741u /* "AGAC" */ >> (k * 8));
            else if (k < 8) v = (uint8_t)(audio_chunk >> ((k - 4) * 8));
            else {
                const uint64_t idx = (uint64_t)(first_frame_count + f) * audio_chunk + (k - 8);
                v = idx < audio_size ? audio[idx] : 0;
            }
        }
        dst[b] = v;
    }
}

// Host driver. bs: batch bitstream (n bytes + >= 64 bytes of readable padding);
// fs: device array of F+1 frame starts (fs[0]=0, fs[F]=n); wk.segs / wk.seg_len
// describe the same frames for the parse and wk.items the hash-link ranges (filled
// by the caller, ntile tiles in total). Results: wk.csize, wk.outbits, wk.chunk_off
// on the device and the chunk image written to `image` (capacity >= 32*F + 9n/8 + 8).
inline void lzss_encode_batch(LzWork& wk, const uint8_t* bs, const uint32_t* fs, uint32_t F, uint32_t n, uint32_t ntile, uint32_t max_usize,
                                    uint32_t first_frame_count, uint8_t* image, LaunchCtx& lc) {
    cudaStream_t st = lc.st;
    KL(lc, KC_LZ_INIT, (lzc_wbase_k<<<cdiv(F + 1, 256), 256, 0, st>>>(fs, F, wk.wbase)));
    const uint32_t* lw15 = wk.lw[0];
    if (n > 0) {
        // persistent walking kernels: enough warps to fill the GPU, chunks handed out through one counter per launch
        const uint32_t nb3 = std::min<uint32_t>(cdiv(n, LZC_BCHUNK), wk.link3_blocks), nb = std::min<uint32_t>(cdiv(n, LZC_BCHUNK), wk.level_blocks);
        cudaMemsetAsync(wk.counters, 0, 16 * sizeof(uint32_t), st);
        KL(lc, KC_LZ_LINK, (lzc_cframe_k<<<cdiv(cdiv(n, LZC_WCHUNK), 256), 256, 0, st>>>(fs, F, n, wk.cframe)));
        KL(lc, KC_LZ_LINK, (lzc_hashlink_k<<<wk.n_items, 32, (size_t)4 << wk.hash_bits, st>>>(bs, n, fs, wk.items, wk.lw[1], wk.rsd, wk.hash_bits)));
        KL(lc, KC_LZ_LINK3, (lzc_link3_k<LZC_ROUNDS><<<nb3, LZC_THREADS, 0, st>>>(bs, fs, F, n, wk.lw[1], wk.rsd, wk.lw[0], wk.cframe, wk.counters)));
        static const int refill_min = getenv("AGMVB_LZ_REFILL") ? atoi(getenv("AGMVB_LZ_REFILL")) : 8;
        int cur = 0;
        for (uint32_t L = LZ_MINLEN; L < (uint32_t)LZ_MAXLEN; L++, cur ^= 1)
            KL(lc, KC_LZ_LEVEL, (lzc_level_k<LZC_ROUNDS><<<nb, LZC_THREADS, 0, st>>>(bs, n, L, wk.lw[cur], wk.rsd, wk.lw[cur ^ 1], fs, wk.cframe,
                                                                                     wk.counters + (L - LZ_MINLEN + 1), refill_min)));
        lw15 = wk.lw[cur];
        KL(lc, KC_LZ_LEVEL, (lzc_bestlen_k<<<cdiv(cdiv(n, 4u), 256u), 256, 0, st>>>(lw15, n, wk.bestlen)));
    }
    orbit_prepare<LZ_MAXLEN, LzStep>(wk.bestlen, wk.segs, F, wk.seg_len, ntile, wk.orb, lc, KC_LZ_PARSE);
    if (n > 0 && ntile) {
        size_t words = (((size_t)n * 9) >> 5) + 3 * (size_t)F + 4;
        cudaMemsetAsync(wk.out_words, 0, words * 4, st);
        KL(lc, KC_LZ_PARSE, (lz_emit_mark_k<<<cdiv(ntile, ORB_LANES), ORB_LANES, 0, st>>>(wk.bestlen, wk.segs, F, wk.seg_len, ntile, wk.orb, bs, lw15, wk.wbase,
                                                                                          wk.out_words, wk.list15, wk.cnt15)));
        KL(lc, KC_LZ_PACK, (lz_pack15_k<<<cdiv(ntile, LZC_WARPS * LZ15_GROUP), LZC_THREADS, 0, st>>>(wk.segs, F, ntile, wk.orb, lw15, wk.rsd, wk.wbase, wk.list15,
                                                                                                     wk.cnt15, wk.out_words)));
    }
    KL(lc, KC_LZ_CHUNK, (lz_finalize_k<<<1, 1024, 0, st>>>(F, wk.stub_bytes, wk.orb.final_cum, wk.outbits, wk.csize, wk.chunk_off)));
    dim3 grid(32, F);
    KL(lc, KC_LZ_CHUNK, (lz_write_chunks_k<<<grid, 256, 0, st>>>(fs, wk.csize, wk.chunk_off, wk.wbase, wk.out_words, first_frame_count, wk.stub_bytes, image,
                                                                 wk.audio_chunk, wk.audio, wk.audio_size)));
}

}  // namespace agmvb
