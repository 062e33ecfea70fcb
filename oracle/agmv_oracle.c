/* TEST INFRASTRUCTURE ONLY - see agmv_oracle.h.
 *
 * Plain-C restatement of the reference's per-frame encode/decode path. Every
 * function cites the reference file:line it follows (paths relative to the
 * reference checkout). Written from the algorithm's description, not copied:
 * the reference streams through FILE* and 8-byte pixels; this works on memory
 * buffers and 4-byte pixels.
 */
#include "agmv_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* synthetic input                                      */
/* ------------------------------------------------------------------ */
static uint32_t mix32(uint32_t h) {
    h ^= h >> 13;
    h *= 0x5bd1e995u;
    h ^= h >> 15;
    return h;
}

void orc_synth_frame(int w, int h, int t, uint32_t seed, uint32_t* out) {
    int s = w / 8 > 8 ? w / 8 : 8;
    int mx = w - s > 1 ? w - s : 1, my = h - s > 1 ? h - s : 1;
    int sx = (5 * t) % mx, sy = (3 * t) % my;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            uint32_t r, g, b;
            if (x >= sx && x < sx + s && y >= sy && y < sy + s) {
                uint32_t nh = mix32(((uint32_t)x * 73856093u) ^ ((uint32_t)y * 19349663u) ^ ((uint32_t)t * 83492791u) ^ seed);
                r = nh & 255; g = (nh >> 8) & 255; b = (nh >> 16) & 255;
            } else if (y >= h / 3 && y < 2 * h / 3) {
                r = (uint32_t)(x * 255 / (w - 1) + 2 * t) & 255;
                g = (uint32_t)(y * 255 / (h - 1)) & 255;
                b = (uint32_t)(x + y + 4 * t) & 255;
            } else {
                uint32_t th = mix32(((uint32_t)(x / 32) * 73856093u) ^ ((uint32_t)(y / 32) * 19349663u) ^ seed);
                r = th & 255; g = (th >> 8) & 255; b = (th >> 16) & 255;
            }
            out[(size_t)y * w + x] = r << 16 | g << 8 | b;
        }
    }
}

static void put32(uint8_t* p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static void put16(uint8_t* p, uint32_t v) { p[0] = v; p[1] = v >> 8; }
static uint32_t get32(const uint8_t* p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24; }

/* 24-bpp uncompressed BMP, rows stored in buffer order (no flip), BGR. This
 * is the layout AGIDL_LoadBMP hands to the encoder for positive heights
 * (extern/agidl/src/agidl_img_bmp.c:973-1002) and the one AGIDL_QuickExport
 * writes (extern/agidl/src/agidl_img_export.c:20-41). */
int orc_write_bmp24(const char* path, int w, int h, const uint32_t* px) {
    size_t stride = ((size_t)w * 3 + 3) & ~(size_t)3;
    size_t img = stride * h;
    uint8_t hdr[54];
    memset(hdr, 0, sizeof hdr);
    hdr[0] = 'B'; hdr[1] = 'M';
    put32(hdr + 2, (uint32_t)(54 + img));
    put32(hdr + 10, 54);
    put32(hdr + 14, 40);
    put32(hdr + 18, (uint32_t)w);
    put32(hdr + 22, (uint32_t)h);
    put16(hdr + 26, 1);
    put16(hdr + 28, 24);
    put32(hdr + 34, (uint32_t)img);
    put32(hdr + 38, 2835);
    put32(hdr + 42, 2835);
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    fwrite(hdr, 1, 54, f);
    uint8_t* row = (uint8_t*)calloc(stride, 1);
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            uint32_t c = px[(size_t)y * w + x];
            row[3 * x] = c & 255; row[3 * x + 1] = (c >> 8) & 255; row[3 * x + 2] = (c >> 16) & 255;
        }
        fwrite(row, 1, stride, f);
    }
    free(row);
    fclose(f);
    return 0;
}

int orc_read_bmp24(const char* path, int* w, int* h, uint32_t* px, size_t cap_px) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    uint8_t hdr[54];
    if (fread(hdr, 1, 54, f) != 54 || hdr[0] != 'B' || hdr[1] != 'M') { fclose(f); return -2; }
    int ww = (int)get32(hdr + 18), hh = (int)get32(hdr + 22);
    if (hdr[28] != 24 || get32(hdr + 30) != 0 || hh <= 0) { fclose(f); return -3; }
    *w = ww; *h = hh;
    if ((size_t)ww * hh > cap_px) { fclose(f); return -4; }
    fseek(f, (long)get32(hdr + 10), SEEK_SET);
    size_t stride = ((size_t)ww * 3 + 3) & ~(size_t)3;
    uint8_t* row = (uint8_t*)malloc(stride);
    for (int y = 0; y < hh; y++) {
        if (fread(row, 1, stride, f) != stride) { free(row); fclose(f); return -5; }
        for (int x = 0; x < ww; x++)
            px[(size_t)y * ww + x] = (uint32_t)row[3 * x + 2] << 16 | (uint32_t)row[3 * x + 1] << 8 | row[3 * x];
    }
    free(row);
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------ */
/* colour math                                                         */
/* ------------------------------------------------------------------ */
static int cr(uint32_t c) { return (c >> 16) & 255; }
static int cg(uint32_t c) { return (c >> 8) & 255; }
static int cb(uint32_t c) { return c & 255; }

/* src/agmv_encode.c:2278-2291 */
uint32_t orc_max_clr(int quality) {
    if (quality == ORC_MID_QUALITY) return 131071;
    if (quality == ORC_LOW_QUALITY) return 65535;
    return 524287;
}

/* src/agmv_utils.c:695-742 */
uint32_t orc_quantize_color(uint32_t c, int q) {
    uint32_t r = cr(c), g = cg(c), b = cb(c);
    if (q == ORC_MID_QUALITY) return (r >> 3) << 12 | (g >> 2) << 6 | (b >> 2);
    if (q == ORC_LOW_QUALITY) return (r >> 3) << 11 | (g >> 2) << 5 | (b >> 3);
    return (r >> 2) << 13 | (g >> 2) << 7 | (b >> 1);
}

/* src/agmv_utils.c:644-693: components of a quantised colour */
static void qparts(uint32_t c, int q, int* r, int* g, int* b) {
    if (q == ORC_MID_QUALITY) { *r = (c & 0x1f000) >> 12; *g = (c & 0xfc0) >> 6; *b = c & 0x3f; }
    else if (q == ORC_LOW_QUALITY) { *r = (c & 0xf800) >> 11; *g = (c & 0x7e0) >> 5; *b = c & 0x1f; }
    else { *r = (c & 0x7e000) >> 13; *g = (c & 0x1f80) >> 7; *b = c & 0x7f; }
}

/* src/agmv_utils.c:744-783 */
static uint32_t dequantize(uint32_t c, int q) {
    int r, g, b;
    qparts(c, q, &r, &g, &b);
    if (q == ORC_MID_QUALITY) { r <<= 3; g <<= 2; b <<= 2; }
    else if (q == ORC_LOW_QUALITY) { r <<= 3; g <<= 2; b <<= 3; }
    else { r <<= 2; g <<= 2; b <<= 1; }
    return (uint32_t)r << 16 | (uint32_t)g << 8 | (uint32_t)b;
}

/* src/agmv_encode.c:554 / :37 - single-palette profiles */
int orc_opt_is_dual(int opt) { return !(opt == ORC_OPT_II || opt == ORC_OPT_ANIM || opt == ORC_OPT_GBA_II); }
/* src/agmv_encode.c:2727 - LIGHT PDIFS for every profile except I, ANIM, GBA_I, GBA_II (BMP input) */
int orc_opt_is_light(int opt) { return !(opt == ORC_OPT_I || opt == ORC_OPT_ANIM || opt == ORC_OPT_GBA_I || opt == ORC_OPT_GBA_II); }

/* src/agmv_encode.c:2389-2394; bin max_clr is out of bounds in the reference
 * and never selectable, so it is not counted. */
void orc_histogram_add(uint64_t* hist, const uint32_t* px, size_t n, int quality) {
    uint32_t mc = orc_max_clr(quality);
    for (size_t i = 0; i < n; i++) {
        uint32_t hcol = orc_quantize_color(px[i], quality);
        if (hcol < mc) hist[hcol]++;
    }
}

typedef struct { uint64_t count; uint32_t idx; } hent;
static int hent_cmp(const void* a, const void* b) {
    const hent* x = (const hent*)a; const hent* y = (const hent*)b;
    if (x->count != y->count) return x->count < y->count ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* src/agmv_utils.c:995-1010 (stable ascending sort by count) +
 * src/agmv_encode.c:2572-2656 (greedy pick, split, de-quantise). */
void orc_build_palette(const uint64_t* hist, int quality, int opt, uint32_t pal0[256], uint32_t pal1[256]) {
    uint32_t mc = orc_max_clr(quality);
    hent* s = (hent*)malloc(sizeof(hent) * mc);
    for (uint32_t i = 0; i < mc; i++) { s[i].count = hist[i]; s[i].idx = i; }
    qsort(s, mc, sizeof(hent), hent_cmp); /* (count, idx) ascending == stable bubble sort */

    uint32_t pal[512];
    memset(pal, 0, sizeof pal);
    memset(pal0, 0, 256 * sizeof(uint32_t));
    memset(pal1, 0, 256 * sizeof(uint32_t));
    int count = 0;
    for (uint32_t n = mc; n > 0; n--) {
        uint32_t clr = n == mc ? 0 : s[n].idx; /* colorgram[max_clr] is an OOB read of a zero page */
        int r, g, b, skip = 0;
        qparts(clr, quality, &r, &g, &b);
        for (int j = 0; j < 512; j++) {
            int pr, pg, pb;
            qparts(pal[j], quality, &pr, &pg, &pb);
            int dr = abs(r - pr), dg = abs(g - pg), db = abs(b - pb);
            if (quality == ORC_HIGH_QUALITY) { if (dr <= 2 && dg <= 2 && db <= 3) skip = 1; }
            else if (dr <= 1 && dg <= 1 && db <= 1) skip = 1;
        }
        if (!skip) pal[count++] = clr;
        if (count >= 512) break;
    }
    free(s);

    if (orc_opt_is_dual(opt)) {
        for (int n = 0; n < 512; n++) {
            uint32_t inv = dequantize(pal[n], quality);
            if (n < 126) pal0[n] = inv;
            else if (n <= 252) pal1[n - 126] = inv;
            if (n > 252 && n <= 381) pal0[n - 126] = inv;
            if (n > 381 && (n - 255) < 256) pal1[n - 255] = inv;
        }
    } else {
        for (int n = 0; n < 256; n++) pal0[n] = dequantize(pal[n], quality);
    }
}

/* src/agmv_utils.c:949-969 */
void orc_interp(uint32_t* dst, const uint32_t* a, const uint32_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) {
        int r1 = cr(a[i]), g1 = cg(a[i]), b1 = cb(a[i]);
        int r2 = cr(b[i]), g2 = cg(b[i]), b2 = cb(b[i]);
        int r = r1 + ((r2 - r1) >> 1), g = g1 + ((g2 - g1) >> 1), bb = b1 + ((b2 - b1) >> 1);
        dst[i] = (uint32_t)(r << 16 | g << 8 | bb);
    }
}

/* extern/agidl/src/agidl_imgp_scale.c:262-293 (float32 index math) */
void orc_scale_nearest(const uint32_t* src, int w, int h, float sx, float sy, uint32_t* dst, int* nw, int* nh) {
    unsigned long worg = (unsigned long)w, horg = (unsigned long)h;
    unsigned long neww = (unsigned long)(worg * sx), newh = (unsigned long)(horg * sy);
    float xscale = (float)(worg - 1) / neww;
    float yscale = (float)(horg - 1) / newh;
    for (unsigned long y = 0; y < newh; y++)
        for (unsigned long x = 0; x < neww; x++) {
            unsigned long x2 = (unsigned long)(x * xscale);
            unsigned long y2 = (unsigned long)(y * yscale);
            dst[y * neww + x] = src[y2 * worg + x2];
        }
    *nw = (int)neww; *nh = (int)newh;
}

/* src/agmv_utils.c:785-816: strict '<' so the lowest index wins ties */
static int nearest_index(const uint32_t* pal, uint32_t c, uint32_t* dist_out) {
    int r = cr(c), g = cg(c), b = cb(c), index = 0;
    uint32_t min = 255 * 255 * 3 + 1;
    for (int i = 0; i < 256; i++) {
        int dr = r - cr(pal[i]), dg = g - cg(pal[i]), db = b - cb(pal[i]);
        uint32_t d = (uint32_t)(dr * dr + dg * dg + db * db);
        if (d < min) { min = d; index = i; }
    }
    *dist_out = min;
    return index;
}

/* src/agmv_utils.c:851-895 (dual: tie goes to palette 0); :589-592 (single) */
uint16_t orc_nearest_entry(const uint32_t* pal0, const uint32_t* pal1, int dual, uint32_t c) {
    uint32_t d0, d1;
    int i0 = nearest_index(pal0, c, &d0);
    if (!dual) return (uint16_t)i0;
    int i1 = nearest_index(pal1, c, &d1);
    return d0 <= d1 ? (uint16_t)i0 : (uint16_t)(256 | i1);
}

void orc_quantize_frame(const uint32_t* px, size_t n, const uint32_t* pal0, const uint32_t* pal1, int dual, uint16_t* entries) {
    /* memoise per distinct colour: the reference recomputes every pixel */
    uint16_t* lut = (uint16_t*)malloc((size_t)1 << 25);
    memset(lut, 0xff, (size_t)1 << 25);
    for (size_t i = 0; i < n; i++) {
        uint32_t c = px[i] & 0xffffff;
        if (lut[c] == 0xffff) lut[c] = orc_nearest_entry(pal0, pal1, dual, c);
        entries[i] = lut[c];
    }
    free(lut);
}

static uint32_t ent_color(uint16_t e, const uint32_t* pal0, const uint32_t* pal1) {
    return (e >> 8) ? pal1[e & 255] : pal0[e & 255];
}
static int within2(uint32_t a, uint32_t b) {
    return abs(cr(a) - cr(b)) <= 2 && abs(cg(a) - cg(b)) <= 2 && abs(cb(a) - cb(b)) <= 2;
}
static size_t put_code(uint8_t* out, size_t pos, uint16_t e, int dual) {
    int pal = e >> 8, idx = e & 255;
    if (!dual) { out[pos++] = (uint8_t)idx; return pos; }
    if (idx < 127) out[pos++] = (uint8_t)(pal << 7 | idx);
    else { out[pos++] = (uint8_t)(pal << 7 | 127); out[pos++] = (uint8_t)idx; }
    return pos;
}

/* src/agmv_encode.c:240-527 */
size_t orc_assemble(const uint16_t* ent, const uint16_t* ient, int w, int h, int is_iframe, int dual,
                    const uint32_t* pal0, const uint32_t* pal1, uint8_t* out) {
    size_t pos = 0;
    for (int y = 0; y < h; y += 4)
        for (int x = 0; x < w; x += 4) {
            uint16_t e0 = ent[(size_t)y * w + x];
            uint32_t color = ent_color(e0, pal0, pal1);
            int c_fill = 0, c_copy = 0;
            for (int j = 0; j < 4; j++)
                for (int i = 0; i < 4; i++) {
                    size_t k = (size_t)(y + j) * w + (x + i);
                    uint32_t bc = ent_color(ent[k], pal0, pal1);
                    if (within2(color, bc)) c_fill++;
                    if (!is_iframe && within2(bc, ent_color(ient[k], pal0, pal1))) c_copy++;
                }
            if (!is_iframe && c_copy >= 13) out[pos++] = 0x5E;
            else if (c_fill >= 14) { out[pos++] = 0x4E; pos = put_code(out, pos, e0, dual); }
            else {
                out[pos++] = 0x2F;
                for (int j = 0; j < 4; j++)
                    for (int i = 0; i < 4; i++) pos = put_code(out, pos, ent[(size_t)(y + j) * w + (x + i)], dual);
            }
        }
    return pos;
}

/* LSB-first bit writer, src/agmv_utils.c:86-112 */
typedef struct { uint8_t* out; size_t n; uint64_t buf; int bits; } bitw;
static void bw_put(bitw* w, uint32_t v, int nb) {
    w->buf |= (uint64_t)v << w->bits;
    w->bits += nb;
    while (w->bits >= 8) { w->out[w->n++] = (uint8_t)w->buf; w->buf >>= 8; w->bits -= 8; }
}
static void bw_flush(bitw* w) { if (w->bits > 0) { w->out[w->n++] = (uint8_t)w->buf; w->buf = 0; w->bits = 0; } }

/* src/agmv_encode.c:106-177: brute-force longest match, earliest start among
 * the longest, window 65535, length 3..15, overlap allowed. */
uint32_t orc_lzss(const uint8_t* in, size_t n, uint8_t* out, size_t* nbytes, uint64_t* outbits_out) {
    bitw w = {out, 0, 0, 0};
    long pos = (long)n, outbits = 0;
    for (long i = 0; i < pos;) {
        int val = in[i];
        long max = 15;
        if (i + max > pos) max = pos - i;
        long start = i - 65535;
        if (start < 0) start = 0;
        long bestlen = 0, beststart = 0;
        while (start < i) {
            const uint8_t* p = (const uint8_t*)memchr(in + start, val, (size_t)(i - start));
            if (!p) break;
            start = p - in;
            long j = 0;
            while (j < max && in[start + j] == in[i + j]) j++;
            if (j > bestlen) { bestlen = j; beststart = start; if (j == max) break; }
            start++;
        }
        if (bestlen < 3) {
            bestlen = 1;
            bw_put(&w, 1, 1);
            bw_put(&w, (uint32_t)val, 8);
            outbits += 9;
        } else {
            bw_put(&w, 0, 1);
            bw_put(&w, (uint32_t)(i - beststart), 16);
            bw_put(&w, (uint32_t)bestlen, 4);
            outbits += 21;
        }
        i += bestlen;
    }
    bw_flush(&w);
    *nbytes = w.n;
    if (outbits_out) *outbits_out = (uint64_t)outbits;
    return (uint32_t)((float)outbits / 8.0f); /* float on purpose: src/agmv_encode.c:176 */
}

/* src/agmv_encode.c:179-238. `stale` = the byte the reference would read at
 * data[pos] when a match runs to the end of the buffer. */
static uint32_t lz77_impl(const uint8_t* in, size_t n, uint8_t* out, size_t* nbytes) {
    long pos = (long)n, outbits = 0;
    size_t o = 0;
    for (long i = 0; i < pos;) {
        int val = in[i];
        long max = 255;
        if (i + max > pos) max = pos - i;
        long start = i - 65535;
        if (start < 0) start = 0;
        long bestlen = 0, bestoff = 0;
        for (; start < i; start++) {
            if (in[start] != val) continue;
            long j = 0;
            while (j < max && in[start + j] == in[i + j]) j++;
            if (j > bestlen) { bestlen = j; bestoff = i - start; }
        }
        if (bestlen > 0) {
            out[o++] = (uint8_t)bestoff; out[o++] = (uint8_t)(bestoff >> 8);
            out[o++] = (uint8_t)bestlen;
            out[o++] = in[i + bestlen]; /* caller pads the input with the stale byte */
            i += bestlen + 1;
        } else {
            out[o++] = 0; out[o++] = 0; out[o++] = 0; out[o++] = in[i];
            i++;
        }
        outbits += 32;
    }
    *nbytes = o;
    return (uint32_t)((float)outbits / 8.0f);
}
uint32_t orc_lz77(const uint8_t* in, size_t n, uint8_t* out, size_t* nbytes) { return lz77_impl(in, n, out, nbytes); }

/* ------------------------------------------------------------------ */
/* whole-sequence encode: src/agmv_encode.c:2270-3657 (BMP case)        */
/* ------------------------------------------------------------------ */
/* src/agmv_utils.c:920-947: fraction of pixels whose truncated float grey value is equal */
static float frame_similarity(const uint32_t* a, const uint32_t* b, size_t n) {
    unsigned long count = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t g1 = (uint8_t)((cr(a[i]) + cg(a[i]) + cb(a[i])) / 3.0f);
        uint8_t g2 = (uint8_t)((cr(b[i]) + cg(b[i]) + cb(b[i])) / 3.0f);
        if (g1 == g2) count++;
    }
    return count / (float)n;
}

enum { MODE_AGMV = 0, MODE_VIDEO = 1, MODE_FULL = 2 };

static long encode_impl(int mode, const uint32_t* frames, int n_src, int w, int h, uint32_t create_n, uint32_t fps,
                        int opt, int quality, int compression, uint8_t* out, size_t out_cap) {
    int dual = orc_opt_is_dual(opt), light = orc_opt_is_light(opt);
    uint32_t mc = orc_max_clr(quality);
    size_t src_px = (size_t)w * h;
    uint32_t start = 1, end = (uint32_t)n_src; /* numbering is irrelevant; only end-start matters */
    uint32_t adjusted = end - start;
    int cw = w, ch = h;
    switch (opt) { /* :2296-2353 */
        case ORC_OPT_I: case ORC_OPT_ANIM: adjusted /= 2; break;
        case ORC_OPT_II: case ORC_OPT_III: adjusted = (uint32_t)(adjusted * 0.75); break;
        case ORC_OPT_GBA_I: case ORC_OPT_GBA_II: cw = 120; ch = 80; adjusted /= 2; break;
        case ORC_OPT_GBA_III: cw = 120; ch = 80; adjusted = (uint32_t)(adjusted * 0.75f); break;
        case ORC_OPT_NDS: cw = 128; ch = 96; adjusted = (uint32_t)(adjusted * 0.75); break;
        default: return -1;
    }
    size_t px = (size_t)cw * ch;

    /* pass 1: histogram over every source frame, unscaled (:2371-2568) */
    uint64_t* hist = (uint64_t*)calloc(mc, sizeof(uint64_t));
    for (int f = 0; f < n_src; f++) orc_histogram_add(hist, frames + (size_t)f * src_px, src_px, quality);
    uint32_t pal0[256], pal1[256];
    orc_build_palette(hist, quality, opt, pal0, pal1);
    free(hist);

    /* header (:21-94) */
    size_t o = 0;
    if (out_cap < 38 + 1536) return -2;
    memset(out, 0, out_cap);
    memcpy(out, "AGMV", 4);
    put32(out + 4, create_n); put32(out + 8, (uint32_t)cw); put32(out + 12, (uint32_t)ch);
    out[16] = 1;
    out[17] = (uint8_t)((compression == ORC_LZSS ? 0 : 2) + (dual ? 1 : 2)); /* src/agmv_utils.c:487-545 */
    put32(out + 18, fps);
    /* audio duration, sample rate, audio size, channels = 0; bits per sample = 16 (CreateAGMV, src/agmv_utils.c:358-366) */
    put16(out + 36, 16);
    o = 38;
    for (int i = 0; i < 256; i++) { out[o++] = cr(pal0[i]); out[o++] = cg(pal0[i]); out[o++] = cb(pal0[i]); }
    if (dual) for (int i = 0; i < 256; i++) { out[o++] = cr(pal1[i]); out[o++] = cg(pal1[i]); out[o++] = cb(pal1[i]); }

    uint16_t* ent = (uint16_t*)malloc(px * 2);
    uint16_t* ient = (uint16_t*)calloc(px, 2);
    uint32_t* tmp = (uint32_t*)malloc(px * 4);
    uint32_t* sa = (uint32_t*)malloc(px * 4);
    uint32_t* sb = (uint32_t*)malloc(px * 4);
    uint8_t* bs = (uint8_t*)calloc(px * 33 / 16 + 64, 1); /* reference buffer is 2*px; see SURVEY 8a E9 */
    uint8_t* lz = (uint8_t*)malloc((px * 33 / 16 + 64) * 4 + 16);
    uint32_t frame_count = 0, encoded = 0;
    long rc = 0;
    int scaled = (cw != w || ch != h);
    float fsx = (float)cw / (unsigned long)w + 0.001f, fsy = (float)ch / (unsigned long)h + 0.001f;
    /* AGMV_EncodeVideo (src/agmv_encode.c:719-2268): leniency per profile (:741-790), groups gated by frame similarity */
    float leniency = (opt == ORC_OPT_II) ? (float)0.1282 : ((opt == ORC_OPT_GBA_I || opt == ORC_OPT_GBA_II || opt == ORC_OPT_GBA_III) ? 0.0f : (float)0.2282);
    uint32_t* sc = (uint32_t*)malloc(px * 4);
    uint32_t* sd = (uint32_t*)malloc(px * 4);

    for (uint32_t i = start; i <= end;) {
        /* the frames this group encodes: LIGHT f(i), interp(f(i+1),f(i+2)), f(i+3); HEAVY interp(f(i),f(i+1)) */
        int plan[3][2], np, step = light ? 4 : 2;
        if (light) { np = 3; plan[0][0] = i; plan[0][1] = -1; plan[1][0] = i + 1; plan[1][1] = i + 2; plan[2][0] = i + 3; plan[2][1] = -1; }
        else { np = 1; plan[0][0] = i; plan[0][1] = i + 1; }
        if (mode == MODE_FULL) { np = 1; plan[0][0] = i; plan[0][1] = -1; step = 1; } /* AGMV_EncodeFullAGMV: every frame as is (:4041-4370) */
        if (mode == MODE_VIDEO) { /* :1153-1190: the pair that would be merged must be similar enough, else one plain frame */
            int pa = light ? i + 1 : i, pb = light ? i + 2 : i + 1;
            if (pb > (int)end) { rc = -3; goto done; }
            const uint32_t* fa = frames + (size_t)(pa - start) * src_px; const uint32_t* fb = frames + (size_t)(pb - start) * src_px;
            if (scaled) { int nw, nh; orc_scale_nearest(fa, w, h, fsx, fsy, sc, &nw, &nh); orc_scale_nearest(fb, w, h, fsx, fsy, sd, &nw, &nh); fa = sc; fb = sd; }
            if (!(frame_similarity(fa, fb, px) >= leniency)) { np = 1; plan[0][0] = i; plan[0][1] = -1; step = 1; }
        }
        for (int p = 0; p < np; p++) {
            const uint32_t* a; const uint32_t* b = NULL;
            if (plan[p][0] > (int)end || plan[p][1] > (int)end) { rc = -3; goto done; }
            a = frames + (size_t)(plan[p][0] - start) * src_px;
            if (plan[p][1] >= 0) b = frames + (size_t)(plan[p][1] - start) * src_px;
            if (scaled) { /* :2707-2721: scale first, interpolate after */
                int nw, nh;
                orc_scale_nearest(a, w, h, fsx, fsy, sa, &nw, &nh);
                if (nw != cw || nh != ch) { rc = -4; goto done; }
                a = sa;
                if (b) { orc_scale_nearest(b, w, h, fsx, fsy, sb, &nw, &nh); b = sb; }
            }
            const uint32_t* img = a;
            if (b) { orc_interp(tmp, a, b, px); img = tmp; }

            /* AGMV_EncodeFrame (:529-634) */
            int is_i = frame_count % 4 == 0;
            orc_quantize_frame(img, px, pal0, pal1, dual, ent);
            size_t usize = orc_assemble(ent, ient, cw, ch, is_i, dual, pal0, pal1, bs);
            size_t nbytes; uint32_t csize;
            if (compression == ORC_LZSS) csize = orc_lzss(bs, usize, lz, &nbytes, NULL);
            else csize = orc_lz77(bs, usize, lz, &nbytes); /* bs[usize] holds the previous frame's byte: stale read */
            if (o + 16 + nbytes + csize + 32 > out_cap) { rc = -2; goto done; }
            memcpy(out + o, "AGFC", 4);
            put32(out + o + 4, frame_count + 1);
            put32(out + o + 8, (uint32_t)usize);
            put32(out + o + 12, csize);
            memcpy(out + o + 16, lz, nbytes);
            o += 16 + csize;                 /* fseek(pos-4), write csize, fseek(csize, CUR) */
            memset(out + o, 0xff, 8); o += 8; /* trailer clobbers the last partial byte */
            if (mode == MODE_AGMV) { memcpy(out + o, "AGAC", 4); put32(out + o + 4, 0); o += 8; } /* empty audio chunk (:707-717); the other two encoders write none */
            if (is_i) memcpy(ient, ent, px * 2);
            frame_count++; encoded++;
        }
        i += step;
        if (mode != MODE_FULL && i + 4 >= end) break; /* :3610-3612, :2220-2222 */
    }
    if (mode == MODE_AGMV) { /* back-patch (:3615-3620) */
        put32(out + 4, encoded);
        float rate = (float)adjusted / (create_n + 1);
        put32(out + 18, (uint32_t)round(fps * rate));
    } else if (mode == MODE_VIDEO) { /* :2223-2230 */
        put32(out + 4, encoded);
        float rate = (float)encoded / create_n;
        put32(out + 18, (uint32_t)round(fps * rate));
    }
    rc = (long)o;
done:
    free(ent); free(ient); free(tmp); free(sa); free(sb); free(bs); free(lz); free(sc); free(sd);
    return rc;
}

long orc_encode_agmv(const uint32_t* frames, int n_src, int w, int h, uint32_t create_n, uint32_t fps,
                     int opt, int quality, int compression, uint8_t* out, size_t out_cap) {
    return encode_impl(MODE_AGMV, frames, n_src, w, h, create_n, fps, opt, quality, compression, out, out_cap);
}
/* AGMV_EncodeVideo creates its own handle: CreateAGMV(end-start, ...) (src/agmv_encode.c:722) */
long orc_encode_video(const uint32_t* frames, int n_src, int w, int h, uint32_t fps, int opt, int quality, int compression,
                      uint8_t* out, size_t out_cap) {
    return encode_impl(MODE_VIDEO, frames, n_src, w, h, (uint32_t)(n_src - 1), fps, opt, quality, compression, out, out_cap);
}
long orc_encode_full(const uint32_t* frames, int n_src, int w, int h, uint32_t create_n, uint32_t fps,
                     int opt, int quality, int compression, uint8_t* out, size_t out_cap) {
    return encode_impl(MODE_FULL, frames, n_src, w, h, create_n, fps, opt, quality, compression, out, out_cap);
}

/* AGMV_EncodeFrame + the empty AGMV_EncodeAudioChunk for n_enc frames with a given palette, starting at
 * agmv->frame_count == first_fc (src/agmv_encode.c:529-634, :707-717). Used by the sharding tests: a rank
 * encodes only its GOP-aligned range of the sequence. No rescaling (coded size == frame size). */
long orc_encode_frames(const uint32_t* frames, int w, int h, const int32_t* src_a, const int32_t* src_b, int n_enc, uint32_t first_fc,
                       const uint32_t* pal0, const uint32_t* pal1, int dual, uint8_t* out, size_t cap) {
    size_t px = (size_t)w * h, o = 0;
    uint16_t* ent = (uint16_t*)malloc(px * 2);
    uint16_t* ient = (uint16_t*)calloc(px, 2);
    uint32_t* tmp = (uint32_t*)malloc(px * 4);
    uint8_t* bs = (uint8_t*)calloc(px * 33 / 16 + 64, 1);
    uint8_t* lz = (uint8_t*)malloc((px * 33 / 16 + 64) * 4 + 16);
    long rc = 0;
    for (int k = 0; k < n_enc; k++) {
        const uint32_t* img = frames + (size_t)src_a[k] * px;
        if (src_b[k] >= 0) { orc_interp(tmp, img, frames + (size_t)src_b[k] * px, px); img = tmp; }
        uint32_t fc = first_fc + (uint32_t)k;
        int is_i = fc % 4 == 0;
        orc_quantize_frame(img, px, pal0, pal1, dual, ent);
        size_t usize = orc_assemble(ent, ient, w, h, is_i, dual, pal0, pal1, bs), nbytes;
        uint32_t csize = orc_lzss(bs, usize, lz, &nbytes, NULL);
        if (o + 32 + nbytes + 8 > cap) { rc = -2; break; }
        memcpy(out + o, "AGFC", 4);
        put32(out + o + 4, fc + 1); put32(out + o + 8, (uint32_t)usize); put32(out + o + 12, csize);
        memcpy(out + o + 16, lz, nbytes);
        o += 16 + csize;
        memset(out + o, 0xff, 8); o += 8;
        memcpy(out + o, "AGAC", 4); put32(out + o + 4, 0); o += 8;
        if (is_i) memcpy(ient, ent, px * 2);
    }
    free(ent); free(ient); free(tmp); free(bs); free(lz);
    return rc ? rc : (long)o;
}

/* ------------------------------------------------------------------ */
/* decode: src/agmv_decode.c:91-410, 527-647                            */
/* ------------------------------------------------------------------ */
typedef struct { const uint8_t* d; size_t len, pos; uint32_t bitbuf; int bitsin; } rd;
static uint8_t rd_byte(rd* r) { return r->pos < r->len ? r->d[r->pos++] : 0; } /* fread past EOF leaves 0 */
/* src/agmv_utils.c:38-59 */
static uint32_t rd_bits(rd* r, int nb) {
    uint32_t v = r->bitbuf >> (8 - r->bitsin);
    while (nb > r->bitsin) { r->bitbuf = rd_byte(r); v |= r->bitbuf << r->bitsin; r->bitsin += 8; }
    r->bitsin -= nb;
    return v & ((1u << nb) - 1);
}
/* src/agmv_utils.c:140-166 */
static void find_next_frame_chunk(rd* r) {
    size_t p = r->pos;
    if (p + 4 <= r->len && !memcmp(r->d + p, "AGFC", 4)) return;
    p += 4;
    while (p < r->len) {
        if (p + 4 <= r->len && !memcmp(r->d + p, "AGFC", 4)) { r->pos = p; return; }
        p++;
    }
    r->pos = r->len; /* not found: the reference ends up just before EOF; decoding then fails */
}

int orc_decode_agmv(const uint8_t* file, size_t len, uint32_t* frames_out, size_t cap_px, int* w, int* h, int* n_frames) {
    if (len < 38 || memcmp(file, "AGMV", 4)) return 1;
    uint32_t nfr = get32(file + 4), W = get32(file + 8), H = get32(file + 12);
    int version = file[17];
    uint32_t fps = get32(file + 18);
    int bps = file[36] | file[37] << 8;
    if (!(version >= 1 && version <= 4) || fps >= 200 || !(bps == 16 || bps == 8)) return 1;
    int dual = version == 1 || version == 3;
    uint32_t pal0[256], pal1[256];
    size_t o = 38;
    memset(pal1, 0, sizeof pal1);
    for (int i = 0; i < 256; i++, o += 3) pal0[i] = (uint32_t)file[o] << 16 | (uint32_t)file[o + 1] << 8 | file[o + 2];
    if (dual) for (int i = 0; i < 256; i++, o += 3) pal1[i] = (uint32_t)file[o] << 16 | (uint32_t)file[o + 1] << 8 | file[o + 2];
    *w = (int)W; *h = (int)H; *n_frames = (int)nfr;
    if (!frames_out) return 0;
    size_t px = (size_t)W * H;
    if (px * nfr > cap_px) return 3;

    uint32_t* img = (uint32_t*)calloc(px, 4);
    uint32_t* ifr = (uint32_t*)calloc(px, 4);
    uint8_t* buf = (uint8_t*)calloc(px * 2 + 64, 1);
    rd r = {file, len, o, 0, 0};
    uint32_t frame_count = 0;
    int rc = 0;
    for (uint32_t f = 0; f < nfr; f++) {
        find_next_frame_chunk(&r);
        if (r.pos + 16 > r.len || memcmp(r.d + r.pos, "AGFC", 4)) { rc = 1; break; }
        uint32_t usize = get32(r.d + r.pos + 8), csize = get32(r.d + r.pos + 12);
        r.pos += 16;
        size_t bpos = 0;
        if (version <= 2) { /* :171-199 */
            uint64_t bits = 0, nbits = (uint64_t)csize * 8;
            while (bits < nbits && bpos < usize) {
                uint32_t flag = rd_bits(&r, 1); bits++;
                if (flag & 1) { buf[bpos++] = (uint8_t)rd_bits(&r, 8); bits += 8; }
                else {
                    uint32_t off = rd_bits(&r, 16), l = rd_bits(&r, 4);
                    bits += 20;
                    size_t p = bpos;
                    for (uint32_t i = 0; i < l; i++) {
                        size_t s = p - off + i; /* unsigned wrap on purpose */
                        if (s < bpos) { buf[bpos] = buf[s]; bpos++; }
                    }
                }
            }
        } else { /* :200-218 */
            for (uint32_t i = 0; i < csize; i += 4) {
                uint32_t off = rd_byte(&r); off |= (uint32_t)rd_byte(&r) << 8;
                uint32_t l = rd_byte(&r); uint8_t lit = rd_byte(&r);
                size_t p = bpos;
                for (uint32_t k = 0; k < l; k++) {
                    size_t s = p - off + k;
                    if (s < bpos) { buf[bpos] = buf[s]; bpos++; }
                }
                buf[bpos++] = lit;
            }
        }
        r.bitbuf = 0; r.bitsin = 0;

        /* block walk (:224-399) */
        size_t bp = 0;
        int esc = 0, invalid = 0;
        for (uint32_t y = 0; y < H && !esc; y += 4)
            for (uint32_t x = 0; x < W && !esc; x += 4) {
                if (bp > bpos) { esc = 1; break; }
                uint8_t fl = buf[bp++];
                while (fl != 0x4E && fl != 0x2F && fl != 0x5E) {
                    fl = buf[bp++];
                    if (bp > bpos) { esc = 1; break; }
                }
                if (fl != 0x4E && fl != 0x2F && fl != 0x5E) invalid = 1;
                if (fl == 0x4E) {
                    uint8_t c = buf[bp++];
                    uint32_t color;
                    if (dual) {
                        const uint32_t* pal = (c & 0x80) ? pal1 : pal0;
                        if ((c & 0x7f) < 127) color = pal[c & 0x7f]; else color = pal[buf[bp++]];
                    } else color = pal0[c];
                    if (x == W - 4 && y == H - 4) color = img[(x - 1) + (size_t)(y + 1) * W];
                    if (bp > bpos) { esc = 1; break; }
                    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) img[(x + i) + (size_t)(y + j) * W] = color;
                } else if (fl == 0x5E) {
                    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) {
                        size_t k = (x + i) + (size_t)(y + j) * W; img[k] = ifr[k];
                    }
                } else {
                    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) {
                        uint8_t c = buf[bp++];
                        uint32_t color;
                        if (dual) {
                            const uint32_t* pal = (c & 0x80) ? pal1 : pal0;
                            if ((c & 0x7f) < 127) color = pal[c & 0x7f]; else color = pal[buf[bp++]];
                            if (bp > bpos || invalid) { esc = 1; invalid = 0; break; }
                        } else {
                            if (bp > bpos || invalid) { esc = 1; invalid = 0; break; }
                            color = pal0[c];
                        }
                        img[(x + i) + (size_t)(y + j) * W] = color;
                    }
                }
            }
        if (frame_count % 4 == 0) memcpy(ifr, img, px * 4);
        frame_count++;
        memcpy(frames_out + (size_t)f * px, img, px * 4);
    }
    free(img); free(ifr); free(buf);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * N4: audio chunk codec
 * ------------------------------------------------------------------------------------------------ */
/* src/agmv_encode.c:15-29 (roundUpEven / roundUpOdd on a u8: 255 wraps to 0) */
static uint8_t up_even(uint8_t v) { while (v % 2 != 0) v++; return v; }
static uint8_t up_odd(uint8_t v) { while (v % 2 == 0) v++; return v; }
/* src/agmv_encode.c:31-36 */
static float orc_round(float x) { return x >= 0.0 ? floor(x + 0.5) : ceil(x - 0.5); }

/* src/agmv_encode.c:666-697. The one conversion the C standard leaves open - (u8)256.0f for samples above 65280, where
 * round(sqrt) is 256 - is written the way the reference's x86-64 build evaluates it (truncate to int, keep the low byte). */
void orc_audio_compress16(const uint16_t* pcm, size_t n, uint8_t* atsample) {
    for (size_t i = 0; i < n; i++) {
        int samp = pcm[i];
        uint8_t ssqrt1 = (uint8_t)sqrt(samp);
        uint8_t ssqrt2 = (uint8_t)(int)orc_round(sqrt(samp));
        uint8_t shift = (uint8_t)(samp >> 8);
        ssqrt1 = up_even(ssqrt1);
        ssqrt2 = up_even(ssqrt2);
        shift = up_odd(shift);
        int resamp1 = ssqrt1 * ssqrt1, resamp2 = ssqrt2 * ssqrt2, resamp3 = shift << 8;
        uint32_t dist1 = abs(resamp1 - samp), dist2 = abs(resamp2 - samp), dist3 = abs(resamp3 - samp);
        uint32_t dist = dist1 < dist3 ? dist1 : dist3; /* :683-684: the min with dist2 is overwritten */
        atsample[i] = dist == dist1 ? ssqrt1 : (dist == dist2 ? ssqrt2 : shift);
    }
}

/* src/agmv_decode.c:433-438 with AGMV_SQR_TABLE[b] = b*b and AGMV_SHIFT_TABLE[b] = 256*b (:21-89) */
void orc_audio_expand16(const uint8_t* atsample, size_t n, uint16_t* pcm) {
    for (size_t i = 0; i < n; i++) {
        unsigned b = atsample[i];
        pcm[i] = (uint16_t)(b % 2 == 0 ? b * b : b * 256);
    }
}
