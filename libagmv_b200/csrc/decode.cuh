// Decode-side kernels D2 (LZSS / LZ77 expand), stale-byte resolution, D3a
// (block record index) and D3b (reconstruction).
// Reference: AGMV_DecodeFrameChunk, src/agmv_decode.c:145-410; bit reader
// src/agmv_utils.c:38-59. SURVEY.md 9.1 is the normative restatement.
//
// The reference decodes into three persistent buffers (pixels, I-frame copy,
// expanded bitstream) and depends on what earlier frames left there (SURVEY
// fact 4): blocks a short frame never reaches keep the previous frame's
// pixels, and the block walk may read up to two bytes past the expanded data,
// i.e. bytes an earlier, longer frame left in the buffer. Both are emulated:
//   * expansion of all frames of a batch is independent (one thread each);
//     `stale_k` then resolves, per frame, the 4 bytes at bpos..bpos+3 from the
//     most recent earlier frame that wrote them (or the carried-over buffer);
//   * reconstruction runs in frame order; every block either decodes from the
//     records or copies the previous frame's block.
#pragma once
#include "common.cuh"

namespace agmvb {

constexpr uint32_t DEC_SLACK = 32;  // a match may overrun usize by up to 14 bytes

struct DecFrame {            // per frame of the batch (host-built, device-read)
    const uint8_t* file;     // the stream's file image on the device
    uint64_t file_len;
    uint64_t data_off;       // file offset of the first payload byte (after the 16-byte chunk header)
    uint64_t ebuf_off;       // offset of this frame's expansion in the batch buffer
    const uint8_t* persist;  // the stream's carried-over expanded-bitstream buffer
    uint32_t persist_len;
    uint32_t usize, csize;
    uint32_t stream_first;   // batch index of the first frame of the same stream
    int lz77;                // stream version 3/4
    int dual;                // stream version 1/3
};

// ---- D2 ----------------------------------------------------------------------
__global__ void __launch_bounds__(64) expand_k(const DecFrame* __restrict__ fr, uint32_t F, uint8_t* __restrict__ ebuf,
                                               uint32_t* __restrict__ bpos_out, uint32_t* __restrict__ consumed_out) {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const DecFrame d = fr[f];
    const uint8_t* __restrict__ file = d.file;
    const uint64_t file_len = d.file_len;
    const int lz77 = d.lz77;
    uint8_t* e = ebuf + d.ebuf_off;
    uint64_t rp = d.data_off;   // next file byte
    uint64_t bpos = 0;
    if (!lz77) {
        uint32_t acc = 0, navail = 0;
        uint64_t bits = 0;
        const uint64_t nbits = (uint64_t)d.csize * 8;
        auto rd = [&](uint32_t nb) -> uint32_t {
            while (navail < nb) {
                uint32_t byte = rp < file_len ? file[rp] : 0u;  // fread past EOF leaves 0
                rp++;
                acc |= byte << navail;
                navail += 8;
            }
            uint32_t v = acc & ((1u << nb) - 1u);
            acc >>= nb;
            navail -= nb;
            return v;
        };
        while (bits < nbits && bpos < d.usize) {
            uint32_t flag = rd(1);
            bits++;
            if (flag) {
                e[bpos++] = (uint8_t)rd(8);
                bits += 8;
            } else {
                uint64_t off = rd(16);
                uint32_t len = rd(4);
                bits += 20;
                const uint64_t p = bpos;
                for (uint32_t i = 0; i < len; i++) {
                    uint64_t s = p - off + i;              // unsigned wrap on purpose: too-large offsets copy nothing
                    if (s < bpos) { e[bpos] = e[s]; bpos++; }
                }
            }
        }
    } else {
        for (uint32_t i = 0; i < d.csize; i += 4) {
            uint32_t b0 = rp < file_len ? file[rp] : 0u; rp++;
            uint32_t b1 = rp < file_len ? file[rp] : 0u; rp++;
            uint32_t len = rp < file_len ? file[rp] : 0u; rp++;
            uint8_t lit = rp < file_len ? file[rp] : 0u; rp++;
            const uint64_t off = b0 | b1 << 8, p = bpos;
            for (uint32_t k = 0; k < len; k++) {
                uint64_t s = p - off + k;
                if (s < bpos) { e[bpos] = e[s]; bpos++; }
            }
            e[bpos++] = lit;
        }
    }
    bpos_out[f] = (uint32_t)bpos;
    consumed_out[f] = (uint32_t)(rp - d.data_off);
}

// ---- stale bytes -----------------------------------------------------------
// stale[f*4+d] = content of the persistent bitstream buffer at index bpos_f + d
// when frame f is walked: written by the latest earlier frame j of the same
// stream with bpos_j > index, else whatever the stream carried in (`persist`).
__global__ void stale_k(const DecFrame* __restrict__ fr, const uint32_t* __restrict__ bpos, uint32_t F,
                        const uint8_t* __restrict__ ebuf, uint8_t* __restrict__ stale) {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int first = (int)fr[f].stream_first;
    for (uint32_t d = 0; d < 4; d++) {
        uint32_t idx = bpos[f] + d;
        uint8_t v = idx < fr[f].persist_len ? fr[f].persist[idx] : 0;
        for (int j = (int)f - 1; j >= first; j--) {
            if (bpos[j] > idx) { v = ebuf[fr[j].ebuf_off + idx]; break; }
        }
        stale[f * 4 + d] = v;
    }
}

// After a batch: fold the batch's expansions into the carried-over buffer.
// frames [first, first+count) of the batch belong to the stream that owns `persist`
__global__ void persist_update_k(const DecFrame* __restrict__ fr, const uint32_t* __restrict__ bpos, uint32_t first, uint32_t count,
                                 const uint8_t* __restrict__ ebuf, uint8_t* __restrict__ persist, uint32_t persist_len) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= persist_len) return;
    for (int j = (int)(first + count) - 1; j >= (int)first; j--) {
        if (bpos[j] > idx) { persist[idx] = ebuf[fr[j].ebuf_off + idx]; return; }
    }
}

// Byte of the virtual persistent buffer as frame f sees it.
struct VBuf {
    const uint8_t* e;
    uint32_t bpos;
    const uint8_t* stale;
    __device__ __forceinline__ uint32_t operator[](uint32_t p) const {
        if (p < bpos) return e[p];
        uint32_t d = p - bpos;
        return d < 4 ? stale[d] : 0u;
    }
};

__device__ __forceinline__ bool is_flag(uint32_t b) { return b == FILL_FLAG || b == NORMAL_FLAG || b == COPY_FLAG; }

// ---- D3a ----------------------------------------------------------------------
// One thread per frame replays the reference's block walk (flags, re-sync
// loop, every `bitpos > bpos` check) and records for each block either
// EMPTY32 (nothing written) or (offset of the first byte after the flag) << 2 | type.
__global__ void __launch_bounds__(64) index_k(const DecFrame* __restrict__ fr, const uint32_t* __restrict__ bpos_arr, uint32_t F,
                                              const uint8_t* __restrict__ ebuf, const uint8_t* __restrict__ stale, uint32_t B,
                                              uint32_t* __restrict__ recs) {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int dual = fr[f].dual;
    const uint32_t bpos = bpos_arr[f];
    const VBuf v{ebuf + fr[f].ebuf_off, bpos, stale + f * 4};
    uint32_t* rec = recs + (size_t)f * B;
    uint32_t bp = 0;
    uint32_t b = 0;
    bool invalid = false;
    for (; b < B; b++) {
        if (bp > bpos) break;
        uint32_t fl = v[bp++];
        bool esc = false;
        while (!is_flag(fl)) {
            fl = v[bp++];
            if (bp > bpos) { esc = true; break; }
        }
        if (!is_flag(fl)) invalid = true;
        if (fl == FILL_FLAG) {
            uint32_t at = bp;
            uint32_t c = v[bp++];
            if (dual && (c & 0x7fu) == 127u) bp++;
            if (bp > bpos) { rec[b] = EMPTY32; b++; break; }
            rec[b] = at << 2 | BT_FILL;
        } else if (fl == COPY_FLAG) {
            rec[b] = bp << 2 | BT_COPY;
        } else {
            uint32_t at = bp;
            bool wrote_any = false;
            for (int k = 0; k < 16; k++) {
                uint32_t c = v[bp++];
                if (dual && (c & 0x7fu) == 127u) bp++;
                if (bp > bpos || invalid) { esc = true; invalid = false; k |= 3; continue; }  // leaves the row, next row re-checks
                wrote_any = true;
            }
            rec[b] = wrote_any ? (at << 2 | BT_NORMAL) : EMPTY32;
        }
        if (esc) { b++; break; }
    }
    for (; b < B; b++) rec[b] = EMPTY32;
}

// ---- D3b ----------------------------------------------------------------------
struct DecStep {              // one frame to reconstruct (host-built)
    uint32_t* dst;
    const uint32_t* prev;     // previous frame's pixels (may alias dst)
    const uint32_t* ifr;      // I-frame pixels
    const uint32_t* recs;
    const uint8_t* ebuf;
    const uint32_t* bpos;
    const uint8_t* stale;
    const uint32_t* pal;      // pal0[256] then pal1[256]
    int dual;
};

__device__ __forceinline__ uint32_t read_color(const VBuf& v, uint32_t& bp, const uint32_t* spal, int dual) {
    uint32_t c = v[bp++];
    if (!dual) return spal[c];
    uint32_t base = (c & 0x80u) ? 256u : 0u;
    if ((c & 0x7fu) < 127u) return spal[base + (c & 0x7fu)];
    return spal[base + v[bp++]];
}

// pixel (i,j) of block `b` as the current frame leaves it
__device__ uint32_t block_pixel(const DecStep& s, const VBuf& v, const uint32_t* spal, uint32_t W, uint32_t b, uint32_t bw, int i, int j) {
    const uint32_t x = (b % bw) * 4 + i, y = (b / bw) * 4 + j;
    const uint32_t r = s.recs[b];
    if (r == EMPTY32) return s.prev[(size_t)y * W + x];
    const uint32_t type = r & 3u;
    uint32_t bp = r >> 2;
    if (type == BT_COPY) return s.ifr[(size_t)y * W + x];
    if (type == BT_FILL) return read_color(v, bp, spal, s.dual);
    uint32_t col = 0;
    for (int k = 0; k <= j * 4 + i; k++) col = read_color(v, bp, spal, s.dual);
    return bp <= v.bpos ? col : s.prev[(size_t)y * W + x];
}

// grid (cdiv(B,128), n_steps): one thread per 4x4 block, 16-byte row stores
__global__ void __launch_bounds__(128) reconstruct_k(const DecStep* __restrict__ steps, uint32_t W, uint32_t H) {
    __shared__ uint32_t spal[512];
    const DecStep s = steps[blockIdx.y];
    for (int k = threadIdx.x; k < 512; k += blockDim.x) spal[k] = s.pal[k];
    __syncthreads();
    const uint32_t bw = W >> 2, B = bw * (H >> 2);
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t x = (b % bw) * 4, y = (b / bw) * 4;
    const uint32_t r = s.recs[b];
    const VBuf v{s.ebuf, *s.bpos, s.stale};
    uint4 row[4];
    if (r == EMPTY32) {
        if (s.dst == s.prev) return;
#pragma unroll
        for (int j = 0; j < 4; j++) row[j] = *reinterpret_cast<const uint4*>(s.prev + (size_t)(y + j) * W + x);
    } else if ((r & 3u) == BT_COPY) {
#pragma unroll
        for (int j = 0; j < 4; j++) row[j] = *reinterpret_cast<const uint4*>(s.ifr + (size_t)(y + j) * W + x);
    } else if ((r & 3u) == BT_FILL) {
        uint32_t bp = r >> 2;
        uint32_t col = read_color(v, bp, spal, s.dual);
        if (b == B - 1) {
            // src/agmv_decode.c:264-266: the last block takes the colour of img[(x-1)+(y+1)*W] as it stands
            // at that moment, i.e. pixel (3,1) of the block to its left as THIS frame leaves it.
            col = bw > 1 ? block_pixel(s, v, spal, W, b - 1, bw, 3, 1) : s.prev[(size_t)(y + 1) * W + x - 1];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) row[j] = make_uint4(col, col, col, col);
    } else {
        uint32_t bp = r >> 2;
        uint32_t px[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t col = read_color(v, bp, spal, s.dual);
            px[k] = bp <= v.bpos ? col : s.prev[(size_t)(y + (k >> 2)) * W + x + (k & 3)];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) row[j] = make_uint4(px[j * 4], px[j * 4 + 1], px[j * 4 + 2], px[j * 4 + 3]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) *reinterpret_cast<uint4*>(s.dst + (size_t)(y + j) * W + x) = row[j];
}

}  // namespace agmvb
