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
#include "orbit.cuh"

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

// ---- D2, warp per frame ------------------------------------------------------
// Token parsing is a serial bit walk (lane 0, payload staged through shared memory 1 KB at a time); the
// copies of up to 32 tokens are then resolved together ("multi-round resolution"): a match may run as soon
// as every byte it reads lies below the output offset of the first unfinished token, so independent matches
// overlap their memory latency instead of paying it one token at a time.
constexpr int EX_WARPS = 4;
constexpr int EX_WIN = 256;  // payload words staged per refill

// LZ77 streams (versions 3 / 4), one warp per frame
__device__ __forceinline__ void expand_lz77_warp(const DecFrame& d, uint8_t* __restrict__ e, const uint8_t* __restrict__ file, uint32_t f, int lane,
                                                 uint32_t* __restrict__ bpos_out, uint32_t* __restrict__ consumed_out) {
    // 4-byte tokens (src/agmv_decode.c:200-218): for (i = 0; i < csize; i += 4) { offset, length, literal }. The tokens are a
    // serial chain, but each one's bytes are independent of each other: byte k of a match comes from
    // p - off + (k mod off), which lies below the token's start p. 32 tokens are fetched per step (one per lane), then
    // every token is copied by the whole warp. Offsets the reference's unsigned arithmetic treats specially (0, or
    // larger than the data so far) take the literal byte-by-byte walk on lane 0.
    // The reference expands into a 2*W*H byte buffer without a bound check; a stream that would run past it is
    // outside its defined behaviour: the expansion stops at the buffer's end here.
    const uint32_t cap = d.persist_len;
    const uint32_t ntok = (d.csize + 3u) / 4u;
    uint32_t bpos = 0, done_tok = 0;
    for (uint32_t t0 = 0; t0 < ntok && bpos < cap; t0 += 32) {
        uint32_t tw = 0;
        if (t0 + lane < ntok) {
            const uint64_t rp = d.data_off + (uint64_t)(t0 + lane) * 4;
#pragma unroll
            for (int j = 0; j < 4; j++) tw |= (rp + j < d.file_len ? (uint32_t)file[rp + j] : 0u) << (8 * j);   // fread past EOF leaves 0
        }
        const uint32_t nb = min(32u, ntok - t0);
        for (uint32_t k = 0; k < nb && bpos < cap; k++) {
            const uint32_t w = __shfl_sync(0xffffffffu, tw, (int)k);
            const uint32_t off = w & 0xFFFFu, len = (w >> 16) & 255u;
            const uint8_t lit = (uint8_t)(w >> 24);
            const uint32_t p = bpos;
            if (off >= 1 && off <= p) {
                const uint32_t nc = min(len, cap - bpos);
                for (uint32_t kk = lane; kk < nc; kk += 32) e[p + kk] = e[p - off + (kk % off)];
                bpos += nc;
            } else if (len > 0 && off != 0) {   // off > p: the first off - p bytes wrap below zero and are skipped
                if (lane == 0) {
                    uint32_t bp = bpos;
                    for (uint32_t kk = 0; kk < len && bp < cap; kk++) {
                        const uint64_t sidx = (uint64_t)p - off + kk;   // unsigned wrap on purpose
                        if (sidx < bp) { e[bp] = e[sidx]; bp++; }
                    }
                    bpos = bp;
                }
                bpos = __shfl_sync(0xffffffffu, bpos, 0);
            }
            if (bpos < cap) { if (lane == 0) e[bpos] = lit; bpos++; }
            done_tok = t0 + k + 1;
            __syncwarp();
        }
    }
    if (lane == 0) {
        bpos_out[f] = bpos;
        consumed_out[f] = done_tok * 4u;
    }
}

__global__ void __launch_bounds__(EX_WARPS * 32) expand_mrr_k(const DecFrame* __restrict__ fr, uint32_t F, uint8_t* __restrict__ ebuf,
                                                             uint32_t* __restrict__ bpos_out, uint32_t* __restrict__ consumed_out) {
    __shared__ uint32_t win[EX_WARPS][EX_WIN + 3];
    __shared__ uint32_t tk_o[EX_WARPS][32];
    __shared__ __align__(16) uint8_t stp[EX_WARPS][EX_WIN * 32];   // per bit of the staged window: length of a token starting there (9 / 21)
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t f = blockIdx.x * EX_WARPS + warp;
    if (f >= F) return;
    const DecFrame d = fr[f];
    uint8_t* e = ebuf + d.ebuf_off;
    const uint8_t* __restrict__ file = d.file;
    if (d.lz77) { expand_lz77_warp(d, e, file, f, lane, bpos_out, consumed_out); return; }
    const uint64_t nbits = (uint64_t)d.csize * 8;
    uint64_t bitp = 0;       // bits consumed so far (== the reference's `bits`), warp-uniform
    uint32_t bpos = 0;       // warp-uniform
    bool more = nbits > 0 && d.usize > 0;
    while (more) {
        // ---- stage the next EX_WIN payload words (plus slack for a token that straddles the end) ----
        const uint64_t win_word0 = bitp >> 5;
        for (int k = lane; k < EX_WIN + 3; k += 32) {
            uint64_t b = d.data_off + (win_word0 + k) * 4;
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) w |= (b + j < d.file_len ? (uint32_t)file[b + j] : 0u) << (8 * j);  // fread past EOF leaves 0
            win[warp][k] = w;
        }
        __syncwarp();
        // Token length for a token starting at every bit of the window (flag bit 1: literal, 9 bits; 0: match, 21 bits), so that
        // the serial walk below is one shared-memory byte load and one add per token. Four bits -> one 32-bit store; lanes write
        // consecutive words.
        for (int it = 0; it < EX_WIN / 4; it++) {
            const int k = it * 4 + (lane >> 3), q = lane & 7;
            const uint32_t nib = (win[warp][k] >> (4 * q)) & 15u;
            reinterpret_cast<uint32_t*>(stp[warp])[k * 8 + q] = 0x15151515u - 12u * ((nib & 1u) | (nib & 2u) << 7 | (nib & 4u) << 14 | (nib & 8u) << 21);
        }
        __syncwarp();
        const uint64_t wbit0 = win_word0 << 5;
        const uint32_t limit = (uint32_t)((nbits - wbit0) < (uint64_t)(EX_WIN * 32) ? (nbits - wbit0) : (uint64_t)(EX_WIN * 32));  // tokens must start below it
        bool window_left = true;
        while (more && window_left) {
            // ---- lane 0: the serial part, reduced to finding where the next <= 32 tokens start (9 or 21 bits each) ----
            uint32_t ntok = 0;
            const uint32_t rel0 = (uint32_t)(bitp - wbit0);
            if (lane == 0) {
                uint32_t rel = rel0;
                const uint8_t* st = stp[warp];
                while (ntok < 32 && rel < limit) {
                    tk_o[warp][ntok++] = rel;
                    rel += st[rel];
                }
            }
            ntok = __shfl_sync(0xffffffffu, ntok, 0);
            __syncwarp();
            if (ntok == 0) { window_left = false; break; }  // next token starts beyond the staged window (or at nbits)
            // ---- all lanes: decode one token each ----
            const bool have = (uint32_t)lane < ntok;
            const uint32_t rel = have ? tk_o[warp][lane] : 0u;
            const uint32_t w = __funnelshift_r(win[warp][rel >> 5], win[warp][(rel >> 5) + 1], rel & 31);
            const bool lit = have && (w & 1u);
            const uint32_t off_raw = (w >> 1) & 0xFFFFu, len_raw = (w >> 17) & 15u;
            uint32_t dd = off_raw;
            uint32_t l = !have ? 0u : (lit ? 1u : len_raw);
            const uint32_t tbits = lit ? 9u : 21u;
            // Output offsets = exclusive prefix of the copy lengths. The reference copies byte i of a match iff
            // (pos - offset + i) < bpos in UNSIGNED arithmetic (src/agmv_decode.c:192-196), so with o = bpos at the token:
            //   1 <= offset <= o          all `len` bytes from o - offset (overlap allowed);
            //   o < offset < o + len      the first (offset - o) indices wrap and are skipped, the rest copy from index 0
            //                             on: len - (offset - o) bytes, i.e. a match of distance o (nothing if o == 0);
            //   otherwise                 nothing.
            // Only damaged streams leave the first case. Lengths are settled front to back: everything before the first
            // lane whose length changes is final.
            uint32_t o;
            while (true) {
                uint32_t inc = l;
#pragma unroll
                for (int k = 1; k < 32; k <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, k); if (lane >= k) inc += y; }
                o = bpos + inc - l;
                uint32_t l2 = l, d2 = dd;
                if (have && !lit) {
                    if (off_raw >= 1 && off_raw <= o) { l2 = len_raw; d2 = off_raw; }
                    else if (off_raw > o && off_raw < o + len_raw && o > 0) { l2 = len_raw - (off_raw - o); d2 = o; }
                    else { l2 = 0; d2 = off_raw; }
                }
                const unsigned bm = __ballot_sync(0xffffffffu, l2 != l || (l2 > 0 && d2 != dd));
                if (!bm) break;
                if (lane == __ffs(bm) - 1) { l = l2; dd = d2; }
            }
            // the reference stops before a token when bpos has reached usize
            const unsigned stop = __ballot_sync(0xffffffffu, have && o >= d.usize);
            const uint32_t nuse = stop ? (uint32_t)(__ffs(stop) - 1) : ntok;
            const bool active = (uint32_t)lane < nuse;
            // new warp-uniform state, taken from the last token actually used
            const uint32_t lastl = nuse - 1;  // nuse >= 1: the first token of a chunk always has o == bpos < usize
            bitp = wbit0 + __shfl_sync(0xffffffffu, rel + tbits, lastl);
            bpos = __shfl_sync(0xffffffffu, o + l, lastl);
            more = bitp < nbits && bpos < d.usize;
            if ((uint32_t)(bitp - wbit0) >= limit && more) window_left = false;
            // ---- resolve the copies of this chunk ----
            bool done = !active || l == 0;
            if (active && lit) { e[o] = (uint8_t)(w >> 1); done = true; }
            __syncwarp();
            while (true) {
                unsigned m = __ballot_sync(0xffffffffu, !done);
                if (!m) break;
                const uint32_t hwm = __shfl_sync(0xffffffffu, o, __ffs(m) - 1);
                const bool can = !done && (o - dd + (l < dd ? l : dd) <= hwm);
                if (can) {
                    const uint8_t* src = e + o - dd;
                    if (dd >= l) {
                        uint8_t v[15];
#pragma unroll
                        for (int k = 0; k < 15; k++) if ((uint32_t)k < l) v[k] = src[k];
#pragma unroll
                        for (int k = 0; k < 15; k++) if ((uint32_t)k < l) e[o + k] = v[k];
                    } else {
                        for (uint32_t k = 0; k < l; k++) e[o + k] = src[k];  // overlapping copy: byte order matters
                    }
                    done = true;
                }
                __syncwarp();
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        bpos_out[f] = bpos;
        consumed_out[f] = (uint32_t)((bitp + 7) >> 3);
    }
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

// ---- D3a, tile-parallel --------------------------------------------------------
// Inside the "safe" prefix of the expanded data (everything more than 72 bytes before bpos) no bounds check of
// the reference's walk can fire, so the walk is a plain pointer chase: flag byte -> record length, non-flag
// byte -> skip one byte (the re-sync loop). index_steps_k turns every byte of the safe prefix into
// (step, counts-as-a-block), the orbit (orbit.cuh) finds the record starts, and index_tail_k finishes the
// last <= 72 bytes with the exact serial logic (index_walk) including stale bytes and escapes.
constexpr uint32_t INDEX_GUARD = 72;

struct IndexStep {
    __device__ static uint32_t step(uint32_t c) { return c & 63u; }
    __device__ static uint32_t weight(uint32_t c) { return c >> 7; }
};

// grid (tiles, F)
__global__ void __launch_bounds__(256) index_steps_k(const DecFrame* __restrict__ fr, const uint32_t* __restrict__ bpos_arr,
                                                     const uint8_t* __restrict__ ebuf, uint8_t* __restrict__ code, uint32_t* __restrict__ seg_len) {
    const uint32_t f = blockIdx.y;
    const uint32_t bpos = bpos_arr[f];
    const uint32_t limit = bpos > INDEX_GUARD ? bpos - INDEX_GUARD : 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) seg_len[f] = limit;
    const int dual = fr[f].dual;
    const uint8_t* e = ebuf + fr[f].ebuf_off;
    uint8_t* c = code + fr[f].ebuf_off;
    const uint32_t lane = lane_id();
    for (uint32_t p0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); p0 < limit; p0 += gridDim.x * blockDim.x) {   // warp-uniform trip count
        const uint32_t p = p0 + lane;
        const uint32_t b = e[p];   // (the expansion is readable 72 bytes past `limit`)
        // A NORMAL record holds sixteen codes of one or two bytes (index 127 escapes to a second byte): its length is a walk
        // over "is an escape code" bits of the 32 bytes behind the flag. Nearly every warp meets a byte 0x2F, so a per-lane
        // loop over e[] ran in every warp with one or two lanes active (82 % issue utilisation,
        // profiles/r02_ncu_summary.txt). Two votes give the bits of this warp's 64 bytes; every lane walks its own window of
        // them in registers, unconditionally.
        uint32_t q = 16;
        if (dual) {
            const unsigned m0 = __ballot_sync(0xffffffffu, (b & 0x7fu) == 127u);
            const unsigned m1 = __ballot_sync(0xffffffffu, (e[p + 32u] & 0x7fu) == 127u);
            const unsigned m = lane == 31u ? m1 : __funnelshift_r(m0, m1, lane + 1u);   // bit i: byte p + 1 + i is an escape code
            q = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) q += 1u + ((m >> q) & 1u);
        }
        uint32_t st = 1, w = 0;
        if (b == COPY_FLAG) w = 1;
        else if (b == FILL_FLAG) { w = 1; st = 2 + ((dual && (e[p + 1] & 0x7fu) == 127u) ? 1u : 0u); }
        else if (b == NORMAL_FLAG) { w = 1; st = 1u + q; }
        if (p < limit) c[p] = (uint8_t)(st | w << 7);
    }
}

struct IndexVisit {
    uint32_t* recs;
    uint32_t B;
    __device__ void operator()(uint32_t sg, uint32_t pos, uint32_t cum, uint32_t cc) const {
        if (!(cc >> 7) || cum >= B) return;
        const uint32_t st = cc & 63u;
        const uint32_t type = st == 1 ? BT_COPY : (st <= 3 ? BT_FILL : BT_NORMAL);
        recs[(size_t)sg * B + cum] = (pos + 1) << 2 | type;
    }
};

// The reference's walk from byte offset bp / block b to the end of the frame (src/agmv_decode.c:224-399).
// keeps[b] (nullable) is set for every block that keeps pixels of the frame before this one: a record the walk never reached or
// left half-way. The reference's csize drops the last partial byte of a frame's payload (src/agmv_encode.c:176), so the last
// block or two of nearly every frame are such blocks.
__device__ inline void index_walk(const VBuf& v, uint32_t bpos, int dual, uint32_t B, uint32_t bp, uint32_t b, uint32_t* __restrict__ rec,
                                  uint8_t* __restrict__ keeps) {
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
            if (bp > bpos) { rec[b] = EMPTY32; if (keeps) keeps[b] = 1; b++; break; }
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
        if ((esc || invalid) && keeps) keeps[b] = 1;
        if (esc) { b++; break; }
    }
    for (; b < B; b++) { rec[b] = EMPTY32; if (keeps) keeps[b] = 1; }
}

__global__ void __launch_bounds__(64) index_tail_k(const DecFrame* __restrict__ fr, const uint32_t* __restrict__ bpos_arr, uint32_t F,
                                                   const uint8_t* __restrict__ ebuf, const uint8_t* __restrict__ stale, uint32_t B,
                                                   const uint32_t* __restrict__ final_pos, const uint32_t* __restrict__ final_cum,
                                                   uint32_t* __restrict__ recs, uint8_t* __restrict__ keeps, uint32_t frames_per_stream) {
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const uint32_t bpos = bpos_arr[f];
    const VBuf v{ebuf + fr[f].ebuf_off, bpos, stale + f * 4};
    uint32_t b0 = final_cum[f];
    // keeps: one byte per (stream, block), set by any frame of the stream (all writers store the same value)
    index_walk(v, bpos, fr[f].dual, B, final_pos[f], b0 < B ? b0 : B, recs + (size_t)f * B, keeps ? keeps + (size_t)(f / frames_per_stream) * B : nullptr);
}

// ---- D3b ----------------------------------------------------------------------
struct DecStep {              // one frame to reconstruct (host side: where its pieces live)
    uint32_t* dst;
    const uint32_t* prev;     // pixels before this frame
    const uint32_t* ifr;      // I-frame snapshot before this frame
    uint32_t fidx;            // the frame's number in the chunk's per-frame arrays (records, bpos, stale bytes)
    uint64_t ebuf_off;        // its expansion inside the chunk's expansion buffer
    int is_snap;              // frame_count % 4 == 0: the frame becomes the I-frame snapshot after it is written
};
// what the kernels read per frame: 16 bytes of position data and the destination, as two arrays [frame of the launch][stream]
struct RecMeta {
    uint64_t ebuf_off;
    uint32_t fidx;
    uint32_t kf;              // bit 31: the frame becomes the snapshot; bits 0-30: the frame's checksum slot
};
struct RecStream {            // per stream of a launch
    const uint32_t* prev;     // pixels before the first frame of the launch
    const uint32_t* ifr;      // I-frame snapshot before the first frame
    const uint32_t* pal;      // pal0[256] then pal1[256]
    unsigned long long* cksum; // nullable: slot k receives the sum over the frame's pixels of value * (2654435761 + 2 * pixel index)
    int dual;
    uint32_t count;           // frames of this stream in the launch
};
struct RecBase { const uint32_t* recs; const uint8_t* ebuf; const uint32_t* bpos; const uint8_t* stale; };
// frames that are not snapshots come in runs of at most three behind their snapshot frame (recon_p_k)
struct RecRun {
    const uint32_t* snap;     // the snapshot the run's COPY blocks show
    const uint32_t* prev;     // pixels before the run's first frame
    uint32_t first, n;        // frames first .. first + n - 1 of the launch's [frame][stream] arrays
    uint32_t sidx, pad;
};

__device__ __forceinline__ uint32_t read_color(const VBuf& v, uint32_t& bp, const uint32_t* spal, int dual) {
    uint32_t c = v[bp++];
    if (!dual) return spal[c];
    uint32_t base = (c & 0x80u) ? 256u : 0u;
    if ((c & 0x7fu) < 127u) return spal[base + (c & 0x7fu)];
    return spal[base + v[bp++]];
}

// pixel number k (row-major 0..15) of a block given its record, the previous value and the snapshot value
__device__ __forceinline__ uint32_t record_pixel(uint32_t r, const VBuf& v, const uint32_t* spal, int dual, int k, uint32_t prev, uint32_t ifr) {
    if (r == EMPTY32) return prev;
    const uint32_t type = r & 3u;
    uint32_t bp = r >> 2;
    if (type == BT_COPY) return ifr;
    if (type == BT_FILL) return read_color(v, bp, spal, dual);
    uint32_t col = 0;
    for (int j = 0; j <= k; j++) col = read_color(v, bp, spal, dual);
    return bp <= v.bpos ? col : prev;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// One block of one frame (src/agmv_decode.c:224-399 for a single block). `ifr`: the snapshot's pixels of the block; `pdst`:
// the previous frame, read back only by a block the frame's walk never reached or left half-way (it keeps those pixels,
// damaged or truncated frames); `fill_last`: the colour a FILL record of the LAST block takes instead of its own
// (img[(x-1)+(y+1)*W] as the frame has left it so far, src/agmv_decode.c:264-266).
__device__ __forceinline__ void paint_block(uint32_t r, const VBuf& v, const uint32_t* spal, int dual, const uint32_t (&ifr)[16], const uint32_t* pdst,
                                            size_t pix0, uint32_t W, bool last, uint32_t fill_last, uint32_t (&cur)[16]) {
    if (r == EMPTY32) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint4 p = *reinterpret_cast<const uint4*>(pdst + pix0 + (size_t)j * W);
            cur[j * 4] = p.x; cur[j * 4 + 1] = p.y; cur[j * 4 + 2] = p.z; cur[j * 4 + 3] = p.w;
        }
    } else if ((r & 3u) == BT_COPY) {
#pragma unroll
        for (int i = 0; i < 16; i++) cur[i] = ifr[i];
    } else if ((r & 3u) == BT_FILL) {
        uint32_t bp = r >> 2;
        uint32_t col = read_color(v, bp, spal, dual);
        if (last) col = fill_last;
#pragma unroll
        for (int i = 0; i < 16; i++) cur[i] = col;
    } else {
        uint32_t bp = r >> 2;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t col = read_color(v, bp, spal, dual);
            cur[i] = bp <= v.bpos ? col : pdst[pix0 + (size_t)(i >> 2) * W + (i & 3)];   // codes past the data: the pixel stays
        }
    }
}
__device__ __forceinline__ void store_block(uint32_t* dst, size_t pix0, uint32_t W, const uint32_t (&cur)[16]) {
#pragma unroll
    for (int j = 0; j < 4; j++) *reinterpret_cast<uint4*>(dst + pix0 + (size_t)j * W) = make_uint4(cur[j * 4], cur[j * 4 + 1], cur[j * 4 + 2], cur[j * 4 + 3]);
}
__device__ __forceinline__ void checksum_block(unsigned long long* slot, bool valid, size_t pix0, uint32_t W, const uint32_t (&cur)[16]) {
    unsigned long long acc = 0;
    if (valid) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const unsigned long long idx = (unsigned long long)pix0 + (unsigned long long)(i >> 2) * W + (i & 3);
            acc += (unsigned long long)cur[i] * (2654435761ull + 2ull * idx);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane_id() == 0) atomicAdd(slot, acc);
}

// The in-order kernel. grid (cdiv(B, 128), n_streams): one thread per 4x4 block position walks frames of its stream in order
// with the block's I-frame snapshot in registers; a frame costs one read of its record and codes and four 16-byte row writes,
// COPY blocks read nothing. A frame is a chain of dependent loads (position data -> record -> codes -> store) and the chain,
// not the bandwidth, sets the pace: 4.7 us per frame and 23 % of DRAM peak whether a thread owns a block or a block row,
// whether the loads of later frames are issued ahead or not (profiles/r02_ncu_summary.txt; a kernel that only writes the
// same pattern runs at 5.9 TB/s, tools/probes/write_pattern.cu). So a thread walks only the frames its block needs in order:
//   * a block that keeps pixels of the previous frame in some frame of the chunk (`keeps`, index_walk), and the last block
//     (its FILL colour follows the neighbour's pixel frame by frame): every frame - list 1;
//   * any other block: the snapshot frames only (frame_count % 4 == 0: a COPY shows the previous snapshot) - list 0; the frames
//     between them are painted by recon_p_k afterwards, all at once.
// keeps == nullptr: every block walks list 1 (ring output, one-block-wide pictures, checksums).
// The position data of frame k+3, the record and bpos of frame k+2 and an L2 prefetch of the codes of frame k+1 are issued
// before frame k is painted.
struct RecList { const RecMeta* meta; uint32_t* const* dsts; const RecStream* streams; };
template <int MINB>
__global__ void __launch_bounds__(128, MINB) reconstruct_k(const RecList snaps, const RecList all, const RecBase base, uint32_t S, uint32_t W, uint32_t H,
                                                           const uint8_t* __restrict__ keeps) {
    __shared__ uint32_t spal[512];
    const uint32_t sidx = blockIdx.y;
    for (int k = threadIdx.x; k < 512; k += blockDim.x) spal[k] = all.streams[sidx].pal[k];
    __syncthreads();
    const uint32_t bw = W >> 2, B = bw * (H >> 2);
    const uint32_t bb = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = bb < B;
    const uint32_t b = valid ? bb : B - 1;  // surplus threads shadow the last block and write nothing
    const bool last = b == B - 1;
    const bool full = !keeps || last || keeps[(size_t)sidx * B + b] != 0;
    const RecList L = full ? all : snaps;
    const RecStream st = L.streams[sidx];
    const uint32_t count = st.count;
    if (count == 0) return;
    const RecMeta* __restrict__ meta = L.meta;
    uint32_t* const* __restrict__ dsts = L.dsts;
    const uint32_t x = (b % bw) * 4, y = (b / bw) * 4;
    const size_t pix0 = (size_t)y * W + x;
    const int dual = st.dual;
    uint32_t ifr[16];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint4 q = *reinterpret_cast<const uint4*>(st.ifr + pix0 + (size_t)j * W);
        ifr[j * 4] = q.x; ifr[j * 4 + 1] = q.y; ifr[j * 4 + 2] = q.z; ifr[j * 4 + 3] = q.w;
    }
    const size_t nb_pix = pix0 + W - 1;   // pixel (x-1, y+1): pixel (3,1) of the left neighbour (one block wide: pixel (3,0) of this block)
    uint32_t nb_ifr = last ? st.ifr[nb_pix] : 0u, nb_prev = last ? st.prev[nb_pix] : 0u;   // (the left neighbour's pixel is another thread's: followed in registers)
    const uint32_t* pdst = st.prev;        // the previous frame's pixels
    auto meta_at = [&](uint32_t k) { return meta[(size_t)(k < count ? k : count - 1) * S + sidx]; };
    RecMeta m0 = meta_at(0), m1 = meta_at(1), m2 = meta_at(2);
    uint32_t r0 = base.recs[(size_t)m0.fidx * B + b], q0 = base.bpos[m0.fidx], n0 = last && bw > 1 ? base.recs[(size_t)m0.fidx * B + b - 1] : 0u;
    uint32_t r1 = base.recs[(size_t)m1.fidx * B + b], q1 = base.bpos[m1.fidx], n1 = last && bw > 1 ? base.recs[(size_t)m1.fidx * B + b - 1] : 0u;
    uint32_t* d0 = dsts[sidx];
    for (uint32_t k = 0; k < count; k++) {
        const RecMeta m3 = meta_at(k + 3);
        const uint32_t r2 = base.recs[(size_t)m2.fidx * B + b], q2 = base.bpos[m2.fidx];
        const uint32_t n2 = last && bw > 1 ? base.recs[(size_t)m2.fidx * B + b - 1] : 0u;
        uint32_t* const d1 = dsts[(size_t)(k + 1 < count ? k + 1 : k) * S + sidx];
        if (r1 != EMPTY32 && (r1 & 3u) != BT_COPY) {
            const uint8_t* c1 = base.ebuf + m1.ebuf_off + (r1 >> 2);
            prefetch_l2(c1);
            if ((r1 & 3u) == BT_NORMAL) prefetch_l2(c1 + 32);
        }
        // ---- frame k ----
        const VBuf v{base.ebuf + m0.ebuf_off, q0, base.stale + (size_t)m0.fidx * 4};
        uint32_t cur[16];
        uint32_t nb_cur = 0;
        if (last) {
            // pixel (x-1, y+1) as this frame has left it when the last block is painted
            if (bw > 1) nb_cur = record_pixel(n0, v, spal, dual, 7, nb_prev, nb_ifr);
            else nb_cur = pdst[nb_pix];   // this block's own pixel (3,0) before the frame
        }
        paint_block(r0, v, spal, dual, ifr, pdst, pix0, W, last, nb_cur, cur);
        if (valid) store_block(d0, pix0, W, cur);
        if (st.cksum) checksum_block(st.cksum + (m0.kf & 0x7fffffffu), valid, pix0, W, cur);   // (only with keeps == nullptr: all lanes walk the same list)
        if (m0.kf >> 31) {
#pragma unroll
            for (int i = 0; i < 16; i++) ifr[i] = cur[i];
            if (last) nb_ifr = nb_cur;
        }
        nb_prev = nb_cur;
        pdst = d0; d0 = d1;
        m0 = m1; m1 = m2; m2 = m3;
        r0 = r1; r1 = r2; q0 = q1; q1 = q2; n0 = n1; n1 = n2;
    }
}

// The frames between snapshots. A frame that is not a snapshot depends on its snapshot frame (COPY blocks) and - damaged or
// truncated frames only - on the frame before it; once the snapshot frames are written (reconstruct_k over those alone), all
// runs of up to three such frames are independent of each other: grid (cdiv(B, 128), runs), a thread paints its block in the
// run's frames, the snapshot's pixels of the block read once if some record of the run is COPY. 1122 of config 3's 1497
// frames become one bandwidth-bound launch instead of 1122 steps of the chain. Blocks on reconstruct_k's list 1 are skipped.
__global__ void __launch_bounds__(128) recon_p_k(const RecRun* __restrict__ runs, const RecMeta* __restrict__ meta, uint32_t* const* __restrict__ dsts,
                                                 const RecStream* __restrict__ streams, const RecBase base, uint32_t S, uint32_t W, uint32_t H,
                                                 const uint8_t* __restrict__ keeps) {
    __shared__ uint32_t spal[512];
    const RecRun run = runs[blockIdx.y];
    const uint32_t sidx = run.sidx;
    const RecStream st = streams[sidx];
    for (int k = threadIdx.x; k < 512; k += blockDim.x) spal[k] = st.pal[k];
    __syncthreads();
    const uint32_t bw = W >> 2, B = bw * (H >> 2);
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B - 1 || keeps[(size_t)sidx * B + b]) return;   // the last block and the blocks that keep previous pixels: reconstruct_k
    const bool valid = true, last = false;
    const uint32_t x = (b % bw) * 4, y = (b / bw) * 4;
    const size_t pix0 = (size_t)y * W + x;
    const int dual = st.dual;
    const size_t nb_pix = pix0 + W - 1;
    RecMeta m[3];
    uint32_t r[3], q[3], nr[3];
    bool any_copy = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const uint32_t k = run.first + min((uint32_t)i, run.n - 1);
        m[i] = meta[(size_t)k * S + sidx];
        r[i] = base.recs[(size_t)m[i].fidx * B + b];
        q[i] = base.bpos[m[i].fidx];
        nr[i] = last && bw > 1 ? base.recs[(size_t)m[i].fidx * B + b - 1] : 0u;
        any_copy = any_copy || (r[i] != EMPTY32 && (r[i] & 3u) == BT_COPY);
    }
    uint32_t ifr[16];
#pragma unroll
    for (int i = 0; i < 16; i++) ifr[i] = 0;
    if (any_copy) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint4 s4 = *reinterpret_cast<const uint4*>(run.snap + pix0 + (size_t)j * W);
            ifr[j * 4] = s4.x; ifr[j * 4 + 1] = s4.y; ifr[j * 4 + 2] = s4.z; ifr[j * 4 + 3] = s4.w;
        }
    }
    const uint32_t nb_ifr = last ? run.snap[nb_pix] : 0u;
    uint32_t nb_prev = last ? run.prev[nb_pix] : 0u;
    const uint32_t* pdst = run.prev;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if ((uint32_t)i < run.n) {
            const VBuf v{base.ebuf + m[i].ebuf_off, q[i], base.stale + (size_t)m[i].fidx * 4};
            uint32_t cur[16];
            uint32_t nb_cur = 0;
            if (last) {
                if (bw > 1) nb_cur = record_pixel(nr[i], v, spal, dual, 7, nb_prev, nb_ifr);
                else nb_cur = pdst[nb_pix];
            }
            paint_block(r[i], v, spal, dual, ifr, pdst, pix0, W, last, nb_cur, cur);
            uint32_t* const d = dsts[(size_t)(run.first + i) * S + sidx];
            if (valid) store_block(d, pix0, W, cur);
            nb_prev = nb_cur;
            pdst = d;
        }
    }
}

}  // namespace agmvb
