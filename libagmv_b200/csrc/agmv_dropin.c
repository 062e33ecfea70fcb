/* libagmv_dropin.so - the reference's C API for the frame hot path (include/agmv_dropin.h)
 * implemented on top of the C-ABI of libagmv_b200.so (include/agmv_b200.h).
 *
 * Host side only: file I/O, handle bookkeeping, 8-byte <-> 4-byte pixel conversion and the
 * container framing the reference does around its per-frame calls. Every pixel-, block- and
 * bitstream-level computation happens in CUDA kernels behind agmvb_*; there is no CPU path
 * (if the GPU layer fails, the encode functions report to stderr and the decode functions
 * return MEMORY_CORRUPTION_ERR, the only channel the reference API offers - SURVEY.md 8b).
 *
 * The audio chunk codec is served too: a handle's track (AGMV_WavToAudioTrack) is companded on the GPU
 * and interleaved by the sequence encoders, AGMV_DecodeAGMV expands the track on the GPU and writes quick_export.wav.
 * Out of scope here: WAV / AIFF import and AIFF export (agmv_utils.c, host file I/O - link the
 * reference's agmv_utils.o for them) and image formats other than BMP (#18, AGIDL).
 */
#include "agmv_dropin.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "agmv_b200.h"

#define WEAK __attribute__((weak))

static agmvb_ctx* g_ctx = NULL;
static int g_device = 0;
static char g_err[600] = "";

const char* AGMV_B200_LastError(void) { return g_err; }
void AGMV_B200_SetDevice(int device) { g_device = device; }

/* AGMV_B200_DEVICES=0,1,2,3: one context per listed GPU for sequence encodes (first use); returns how many there are */
static agmvb_ctx* g_multi[16];
static int g_n_multi = -1;
static int multi_devices(void) {
    if (g_n_multi >= 0) return g_n_multi;
    g_n_multi = 0;
    const char* e = getenv("AGMV_B200_DEVICES");
    if (!e || !*e) return 0;
    while (*e && g_n_multi < 16) {
        char* end = NULL;
        long d = strtol(e, &end, 10);
        if (end == e) break;
        if (agmvb_create(&g_multi[g_n_multi], (int)d, NULL) != 0) {
            fprintf(stderr, "libagmv_dropin: AGMV_B200_DEVICES: no context on GPU %ld, using one GPU\n", d);
            for (int k = 0; k < g_n_multi; k++) agmvb_destroy(g_multi[k]);
            g_n_multi = 0;
            return 0;
        }
        g_n_multi++;
        e = *end == ',' ? end + 1 : end;
    }
    return g_n_multi;
}

static agmvb_ctx* ctx_get(void) {
    if (!g_ctx) {
        int rc = agmvb_create(&g_ctx, g_device, NULL);
        if (rc != AGMVB_OK) {
            snprintf(g_err, sizeof g_err, "agmvb_create failed with %d: no usable CUDA device (this build has no CPU path)", rc);
            fprintf(stderr, "libagmv_dropin: %s\n", g_err);
            g_ctx = NULL;
        }
    }
    return g_ctx;
}

static int fail(int rc, const char* where) {
    snprintf(g_err, sizeof g_err, "%s: error %d: %s", where, rc, g_ctx ? agmvb_last_error(g_ctx) : "no context");
    fprintf(stderr, "libagmv_dropin: %s\n", g_err);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* handle lifecycle (src/agmv_utils.c:332-421), weak: the reference's agmv_utils.o wins     */
/* ------------------------------------------------------------------------------------ */
WEAK AGMV* CreateAGMV(u32 num_of_frames, u32 width, u32 height, u32 frames_per_second) {
    AGMV* a = (AGMV*)malloc(sizeof(AGMV));
    a->frame_chunk = (AGMV_FRAME_CHUNK*)malloc(sizeof(AGMV_FRAME_CHUNK));
    a->audio_chunk = (AGMV_AUDIO_CHUNK*)malloc(sizeof(AGMV_AUDIO_CHUNK));
    a->bitstream = (AGMV_BITSTREAM*)malloc(sizeof(AGMV_BITSTREAM));
    a->bitstream->len = width * height * 2;
    a->bitstream->pos = 0;
    a->bitstream->data = (u8*)malloc(a->bitstream->len);
    a->frame = (AGMV_FRAME*)malloc(sizeof(AGMV_FRAME));
    a->frame->img_data = (u32*)malloc(sizeof(u32) * width * height);
    a->iframe = (AGMV_FRAME*)malloc(sizeof(AGMV_FRAME));
    a->iframe->img_data = (u32*)malloc(sizeof(u32) * width * height);
    a->audio_track = (AGMV_AUDIO_TRACK*)malloc(sizeof(AGMV_AUDIO_TRACK));
    a->iframe_entries = (AGMV_ENTRY*)malloc(sizeof(AGMV_ENTRY) * width * height);
    a->audio_track->pcm = NULL;
    a->audio_track->pcm8 = NULL;
    a->audio_track->start_point = 0;
    a->audio_chunk->atsample = NULL;
    a->frame_count = 0;
    a->header.width = a->frame->width = a->iframe->width = width;
    a->header.height = a->frame->height = a->iframe->height = height;
    a->header.num_of_frames = num_of_frames;
    a->header.frames_per_second = frames_per_second;
    a->header.total_audio_duration = 0;
    a->header.sample_rate = 0;
    a->header.audio_size = 0;
    a->header.num_of_channels = 0;
    a->header.bits_per_sample = 16;
    a->leniency = 0.1282f;
    a->opt = AGMV_OPT_I;
    a->compression = AGMV_LZSS_COMPRESSION;
    a->volume = 1.0f;
    return a;
}

WEAK void DestroyAGMV(AGMV* a) {
    if (!a) return;
    free(a->iframe_entries);
    free(a->frame->img_data);
    free(a->iframe->img_data);
    free(a->frame);
    free(a->iframe);
    free(a->bitstream->data);
    free(a->bitstream);
    free(a->frame_chunk);
    if (a->header.total_audio_duration != 0) {
        if (a->header.bits_per_sample == 16) free(a->audio_track->pcm); else free(a->audio_track->pcm8);
        free(a->audio_chunk->atsample);
    }
    free(a->audio_chunk);
    free(a->audio_track);
    free(a);
}

/* ------------------------------------------------------------------------------------ */
/* small host helpers                                                                      */
/* ------------------------------------------------------------------------------------ */
static void w32(FILE* f, u32 v) { uint32_t x = (uint32_t)v; fwrite(&x, 4, 1, f); }
static void w16(FILE* f, u32 v) { uint16_t x = (uint16_t)v; fwrite(&x, 2, 1, f); }
/* AGMV_ReadFourCC: four fgetc calls, so the end of the file reads as 0xFF bytes and never as the previous chunk's tag */
static void read_fourcc(FILE* f, char* fourcc) { for (int i = 0; i < 4; i++) fourcc[i] = (char)fgetc(f); }
static u32 r32(FILE* f) { uint32_t x = 0; if (fread(&x, 1, 4, f) != 4) { /* short read leaves what arrived */ } return x; }
static u32 r16(FILE* f) { uint16_t x = 0; if (fread(&x, 1, 2, f) != 2) { } return x; }
static u32 r8(FILE* f) { uint8_t x = 0; if (fread(&x, 1, 1, f) != 1) { } return x; }

static int opt_is_dual(AGMV_OPT o) { return !(o == AGMV_OPT_II || o == AGMV_OPT_ANIM || o == AGMV_OPT_GBA_II); }
static int opt_is_light(AGMV_OPT o) { return !(o == AGMV_OPT_I || o == AGMV_OPT_ANIM || o == AGMV_OPT_GBA_I || o == AGMV_OPT_GBA_II); }
/* src/agmv_utils.c:487-545 */
static u8 version_of(AGMV_OPT o, AGMV_COMPRESSION c) { return (u8)((c == AGMV_LZSS_COMPRESSION ? 0 : 2) + (opt_is_dual(o) ? 1 : 2)); }

/* 24/32-bit uncompressed BMP -> 0x00RRGGBB words in file row order (what AGIDL_LoadBMP +
 * AGIDL_ColorConvertBMP(RGB_888) hand the encoder, extern/agidl/src/agidl_img_bmp.c:973-1002). */
static int load_bmp(const char* path, u32 w, u32 h, uint32_t* out) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "libagmv_dropin: cannot open %s\n", path); return -1; }
    uint8_t hd[54];
    if (fread(hd, 1, 54, f) != 54 || hd[0] != 'B' || hd[1] != 'M') { fclose(f); return -2; }
    uint32_t off, bw, bh, bits, comp;
    memcpy(&off, hd + 10, 4); memcpy(&bw, hd + 18, 4); memcpy(&bh, hd + 22, 4);
    bits = hd[28] | hd[29] << 8; memcpy(&comp, hd + 30, 4);
    if (bw != w || bh != h || !(bits == 24 || bits == 32) || comp != 0) {
        fprintf(stderr, "libagmv_dropin: %s: only uncompressed 24/32-bit BMPs of %lux%lu are served\n", path, w, h);
        fclose(f);
        return -3;
    }
    size_t bpp = bits / 8, stride = (w * bpp + 3) & ~(size_t)3;
    uint8_t* row = (uint8_t*)malloc(stride);
    fseek(f, (long)off, SEEK_SET);
    for (u32 y = 0; y < h; y++) {
        if (fread(row, 1, stride, f) != stride) { free(row); fclose(f); return -4; }
        for (u32 x = 0; x < w; x++) out[y * w + x] = (uint32_t)row[x * bpp + 2] << 16 | (uint32_t)row[x * bpp + 1] << 8 | row[x * bpp];
    }
    free(row);
    fclose(f);
    return 0;
}

static unsigned long g_expcount = 0; /* AGIDL's process-global export counter, extern/agidl/src/agidl_img_export.c:18 */

/* AGIDL_QuickExport for BMP / RGB_888 (extern/agidl/src/agidl_img_export.c:20-41, agidl_img_bmp.c:1041-1110):
 * 54-byte header with zero resolution fields, rows in buffer order, BGR. */
static void quick_export_bmp(const uint32_t* px, u32 w, u32 h) {
    char name[64];
    g_expcount++;
    snprintf(name, sizeof name, "quick_export_%lu.bmp", g_expcount);
    FILE* f = fopen(name, "wb");
    if (!f) return;
    uint8_t hd[54];
    memset(hd, 0, sizeof hd);
    uint32_t v;
    hd[0] = 'B'; hd[1] = 'M';
    v = (uint32_t)(54 + w * h * 3); memcpy(hd + 2, &v, 4);
    v = 54; memcpy(hd + 10, &v, 4);
    v = 40; memcpy(hd + 14, &v, 4);
    v = (uint32_t)w; memcpy(hd + 18, &v, 4);
    v = (uint32_t)h; memcpy(hd + 22, &v, 4);
    hd[26] = 1; hd[28] = 24;
    v = (uint32_t)(w * h * 3); memcpy(hd + 34, &v, 4);
    fwrite(hd, 1, 54, f);
    uint8_t* row = (uint8_t*)malloc(w * 3);
    for (u32 y = 0; y < h; y++) {
        for (u32 x = 0; x < w; x++) {
            uint32_t c = px[y * w + x];
            row[3 * x] = (uint8_t)c; row[3 * x + 1] = (uint8_t)(c >> 8); row[3 * x + 2] = (uint8_t)(c >> 16);
        }
        fwrite(row, 1, w * 3, f);
    }
    free(row);
    fclose(f);
}

/* ------------------------------------------------------------------------------------ */
/* encode                                                                                  */
/* ------------------------------------------------------------------------------------ */
/* src/agmv_encode.c:21-94 */
void AGMV_EncodeHeader(FILE* file, AGMV* agmv) {
    fwrite("AGMV", 1, 4, file);
    w32(file, agmv->header.num_of_frames);
    w32(file, agmv->header.width);
    w32(file, agmv->header.height);
    fputc(1, file);
    fputc(version_of(agmv->opt, agmv->compression), file);
    w32(file, agmv->header.frames_per_second);
    w32(file, agmv->header.total_audio_duration);
    w32(file, agmv->header.sample_rate);
    w32(file, agmv->header.audio_size);
    w16(file, agmv->header.num_of_channels);
    w16(file, agmv->header.bits_per_sample);
    for (int p = 0; p < (opt_is_dual(agmv->opt) ? 2 : 1); p++) {
        const u32* pal = p ? agmv->header.palette1 : agmv->header.palette0;
        for (int i = 0; i < 256; i++) { fputc((int)((pal[i] >> 16) & 255), file); fputc((int)((pal[i] >> 8) & 255), file); fputc((int)(pal[i] & 255), file); }
    }
}

/* encoder configuration currently loaded in the GPU context (per-frame API) */
static struct { u32 w, h; int opt, comp, valid; const AGMV* owner; } g_enc = {0, 0, 0, 0, 0, NULL};

static int enc_prepare(agmvb_ctx* c, AGMV* agmv) {
    u32 w = agmv->header.width, h = agmv->header.height;
    /* The state carried from frame to frame (the LZ77 coder's bitstream buffer, the I-frame entries) belongs to ONE handle's
     * sequence: CreateAGMV hands the reference a zeroed buffer. So the GPU context is set up again whenever another handle
     * shows up or a handle starts over at frame 0, not only when the configuration changes. */
    if (!g_enc.valid || g_enc.owner != agmv || agmv->frame_count == 0 || g_enc.w != w || g_enc.h != h || g_enc.opt != (int)agmv->opt ||
        g_enc.comp != (int)agmv->compression) {
        /* coded size == handle size here: scaling (GBA / NDS) is AGMV_EncodeAGMV's job, the per-frame call gets scaled frames */
        int o = agmv->opt;
        if (o == AGMV_OPT_GBA_I || o == AGMV_OPT_GBA_III || o == AGMV_OPT_NDS) o = AGMV_OPT_III; /* same codec path, no rescale */
        if (o == AGMV_OPT_GBA_II) o = AGMV_OPT_II;
        int rc = agmvb_enc_begin(c, (uint32_t)w, (uint32_t)h, o, AGMVB_HIGH_QUALITY, (int)agmv->compression);
        if (rc) return rc;
        g_enc.w = w; g_enc.h = h; g_enc.opt = (int)agmv->opt; g_enc.comp = (int)agmv->compression; g_enc.valid = 1; g_enc.owner = agmv;
    }
    uint32_t p0[256], p1[256];
    for (int i = 0; i < 256; i++) { p0[i] = (uint32_t)agmv->header.palette0[i]; p1[i] = (uint32_t)agmv->header.palette1[i]; }
    return agmvb_enc_set_palette(c, p0, p1);
}

/* src/agmv_encode.c:529-634 */
void AGMV_EncodeFrame(FILE* file, AGMV* agmv, u32* img_data) {
    agmvb_ctx* c = ctx_get();
    if (!c) return;
    const size_t P = (size_t)agmv->header.width * agmv->header.height;
    int rc = enc_prepare(c, agmv);
    if (rc) { fail(rc, "AGMV_EncodeFrame"); return; }
    uint32_t* px = (uint32_t*)malloc(P * 4);
    uint16_t* ent = (uint16_t*)malloc(P * 2);
    for (size_t i = 0; i < P; i++) { px[i] = (uint32_t)img_data[i]; agmv->frame->img_data[i] = img_data[i]; } /* AGMV_SyncFrameAndImage */
    const int is_i = agmv->frame_count % 4 == 0;
    if (!is_i) { /* the I-frame entries live in the caller's handle */
        for (size_t i = 0; i < P; i++) ent[i] = (uint16_t)((agmv->iframe_entries[i].pal_num & 1) << 8 | agmv->iframe_entries[i].index);
        rc = agmvb_enc_set_iframe_entries(c, ent);
    }
    int32_t sa = 0, sb = -1;
    uint64_t nbytes = 0;
    if (!rc) rc = agmvb_enc_set_audio_stub(c, 0); /* the 'AGAC' chunk is AGMV_EncodeAudioChunk's, not ours */
    if (!rc) rc = agmvb_enc_frames(c, px, 1, 0, &sa, &sb, 1, (uint32_t)agmv->frame_count, &nbytes);
    uint8_t* img = (uint8_t*)malloc(nbytes + 16);
    uint32_t us = 0, cs = 0;
    if (!rc) rc = agmvb_enc_fetch(c, img, nbytes, &us, &cs);
    if (!rc) {
        fwrite(img, 1, (size_t)nbytes, file); /* 'AGFC' header, payload, 8 x 0xFF */
        agmv->bitstream->pos = us;
        if (is_i) {
            rc = agmvb_enc_get_iframe_entries(c, ent);
            if (!rc) for (size_t i = 0; i < P; i++) { agmv->iframe_entries[i].pal_num = (u8)(ent[i] >> 8); agmv->iframe_entries[i].index = (u8)ent[i]; }
        }
        agmv->frame_count++;
    }
    if (rc) fail(rc, "AGMV_EncodeFrame");
    free(px); free(ent); free(img);
}

/* the handle's audio track (AGMV_WavToAudioTrack, src/agmv_utils.c:1035-1087) -> GPU layer; AGMV_CompressAudio runs there */
static int track_to_gpu(agmvb_ctx* c, AGMV* agmv) {
    if (agmv->header.audio_size == 0) return agmvb_enc_set_audio(c, NULL, 0, 16, 0, 0, 0);
    const void* pcm = agmv->header.bits_per_sample == 16 ? (const void*)agmv->audio_track->pcm : (const void*)agmv->audio_track->pcm8;
    return agmvb_enc_set_audio(c, pcm, (uint64_t)agmv->header.audio_size, (int)agmv->header.bits_per_sample, (uint32_t)agmv->header.sample_rate,
                               (uint32_t)agmv->header.num_of_channels, (uint32_t)agmv->header.total_audio_duration);
}

/* src/agmv_encode.c:659-705: fills agmv->audio_chunk->atsample (allocated by the caller, audio_size bytes) - on the GPU */
void AGMV_CompressAudio(AGMV* agmv) {
    agmvb_ctx* c = ctx_get();
    if (!c || agmv->header.audio_size == 0) return;
    const void* pcm = agmv->header.bits_per_sample == 16 ? (const void*)agmv->audio_track->pcm : (const void*)agmv->audio_track->pcm8;
    int rc = agmvb_audio_compress(c, pcm, (uint64_t)agmv->header.audio_size, (int)agmv->header.bits_per_sample, agmv->audio_chunk->atsample);
    if (rc) fail(rc, "AGMV_CompressAudio");
}

/* src/agmv_encode.c:707-717 (container framing only) */
void AGMV_EncodeAudioChunk(FILE* file, AGMV* agmv) {
    const u32 size = agmv->audio_chunk->size;
    fwrite("AGAC", 1, 4, file);
    w32(file, size);
    fwrite(agmv->audio_chunk->atsample + agmv->audio_track->start_point, 1, (size_t)size, file);
    agmv->audio_track->start_point += size;
}

static void export_gba_header(const char* filename);

/* src/agmv_encode.c:2270-3657, BMP input */
void AGMV_EncodeAGMV(AGMV* agmv, const char* filename, const char* dir, const char* basename, u8 img_type, u32 start_frame,
                     u32 end_frame, u32 width, u32 height, u32 frames_per_second, AGMV_OPT opt, AGMV_QUALITY quality,
                     AGMV_COMPRESSION compression) {
    (void)frames_per_second; /* the reference writes the handle's rate, not this argument (:41, :3620) */
    agmvb_ctx* c = ctx_get();
    if (!c) { DestroyAGMV(agmv); return; }
    if (img_type != AGMV_IMG_BMP) {
        fprintf(stderr, "libagmv_dropin: AGMV_EncodeAGMV serves BMP input only (image decoding is AGIDL's, out of scope)\n");
        DestroyAGMV(agmv);
        return;
    }
    agmv->opt = opt;
    agmv->compression = compression;
    agmv->leniency = 0;
    u32 adjusted = end_frame - start_frame;
    u32 cw = width, ch = height;
    switch (opt) { /* :2296-2353 */
        case AGMV_OPT_I: case AGMV_OPT_ANIM: adjusted /= 2; break;
        case AGMV_OPT_II: case AGMV_OPT_III: adjusted *= 0.75; break;
        case AGMV_OPT_GBA_I: case AGMV_OPT_GBA_II: cw = 120; ch = 80; adjusted /= 2; break;
        case AGMV_OPT_GBA_III: cw = 120; ch = 80; adjusted *= 0.75f; break;
        case AGMV_OPT_NDS: cw = 128; ch = 96; adjusted *= 0.75; break;
    }
    if (cw != width || ch != height) {
        agmv->header.width = agmv->frame->width = agmv->iframe->width = cw;
        agmv->header.height = agmv->frame->height = agmv->iframe->height = ch;
        free(agmv->frame->img_data);
        agmv->frame->img_data = (u32*)malloc(sizeof(u32) * cw * ch);
    }
    const int light = opt_is_light(opt);
    const size_t SP = (size_t)width * height;
    const u32 n_src = end_frame - start_frame + 1;
    const int cur_dir = dir[0] == 'c' && dir[1] == 'u' && dir[2] == 'r'; /* :2373 */
    char path[512];
    /* AGMV_B200_DEVICES=0,1,..: the sequence is sharded by frame range over those GPUs (agmvb_encode_sequence_multi). The
     * frames are held in host memory for that (both passes read them); sequences with an audio track or the LZ77 coder take the
     * single-GPU path below. The file is byte-identical either way. */
    if (multi_devices() > 1 && compression == AGMV_LZSS_COMPRESSION && agmv->header.audio_size == 0 && n_src >= 4) {
        uint32_t* all = (uint32_t*)malloc((size_t)n_src * SP * 4);
        size_t cap = 4096 + (size_t)n_src * (SP * 3 + 64);
        uint8_t* outb = (uint8_t*)malloc(cap);
        int rcm = (all && outb) ? 0 : AGMVB_ERR_MEMORY;
        for (u32 k = 0; k < n_src && !rcm; k++) {
            if (cur_dir) snprintf(path, sizeof path, "%s%lu.bmp", basename, start_frame + k);
            else snprintf(path, sizeof path, "%s/%s%lu.bmp", dir, basename, start_frame + k);
            if (load_bmp(path, width, height, all + (size_t)k * SP)) rcm = AGMVB_ERR_FILE;
        }
        uint64_t len = 0;
        uint32_t ne = 0;
        if (!rcm) rcm = agmvb_encode_sequence_multi(g_multi, g_n_multi, all, (uint32_t)n_src, (uint32_t)width, (uint32_t)height,
                                                   (uint32_t)agmv->header.num_of_frames, (uint32_t)agmv->header.frames_per_second, (int)opt,
                                                   (int)quality, (int)compression, outb, cap, &len, &ne);
        if (!rcm) {
            FILE* mf = fopen(filename, "wb");
            if (mf) { fwrite(outb, 1, (size_t)len, mf); fclose(mf); }
            else fprintf(stderr, "libagmv_dropin: cannot create %s\n", filename);
        } else {
            snprintf(g_err, sizeof g_err, "AGMV_EncodeAGMV (multi-GPU): error %d: %s", rcm, agmvb_last_error(g_multi[0]));
            fprintf(stderr, "libagmv_dropin: %s\n", g_err);
        }
        free(all); free(outb);
        g_enc.valid = 0;
        DestroyAGMV(agmv);
        if (!rcm && (opt == AGMV_OPT_GBA_I || opt == AGMV_OPT_GBA_II || opt == AGMV_OPT_GBA_III)) export_gba_header(filename);
        return;
    }
    int rc = agmvb_enc_begin(c, (uint32_t)width, (uint32_t)height, (int)opt, (int)quality, (int)compression);
    if (!rc) rc = agmvb_enc_set_audio_stub(c, 1); /* AGMV_EncodeAudioChunk after every frame: 'AGAC' 0 when there is no audio */
    if (!rc) rc = track_to_gpu(c, agmv);          /* AGMV_CompressAudio (:2667) */
    if (!rc && agmv->header.audio_size != 0)      /* audio_chunk->size = audio_size / (f32)adjusted frames (:2661-2663) */
        rc = agmvb_enc_set_audio_chunk(c, (uint32_t)(u32)(agmv->header.audio_size / (f32)adjusted));
    g_enc.valid = 0;
    /* frames are streamed from disk twice, like the reference: once for the histogram, once for the encode */
    const u32 CH = 64; /* source frames per upload; a multiple of 16 keeps LIGHT groups and GOPs whole */
    uint32_t* buf = (uint32_t*)malloc((size_t)(CH + 4) * SP * 4);
    for (u32 f0 = 0; !rc && f0 < n_src; f0 += CH) {
        u32 nf = n_src - f0 < CH ? n_src - f0 : CH;
        for (u32 k = 0; k < nf; k++) {
            if (cur_dir) snprintf(path, sizeof path, "%s%lu.bmp", basename, start_frame + f0 + k);
            else snprintf(path, sizeof path, "%s/%s%lu.bmp", dir, basename, start_frame + f0 + k);
            if (load_bmp(path, width, height, buf + (size_t)k * SP)) { rc = AGMVB_ERR_FILE; break; }
        }
        if (!rc) rc = agmvb_enc_histogram(c, buf, nf, 0);
    }
    if (!rc) rc = agmvb_enc_build_palette(c);
    uint32_t p0[256], p1[256];
    if (!rc) rc = agmvb_enc_get_palette(c, p0, p1);
    if (rc) { fail(rc, "AGMV_EncodeAGMV (palette pass)"); free(buf); DestroyAGMV(agmv); return; }
    for (int i = 0; i < 256; i++) { agmv->header.palette0[i] = p0[i]; agmv->header.palette1[i] = p1[i]; }
    FILE* file = fopen(filename, "wb");
    if (!file) { fprintf(stderr, "libagmv_dropin: cannot create %s\n", filename); free(buf); DestroyAGMV(agmv); return; }
    AGMV_EncodeHeader(file, agmv);

    /* pass 2: PDIFS schedule (:2727-2770), loop exit (:3610-3612); encode in batches of whole groups */
    u32 encoded = 0;
    const u32 step = light ? 4 : 2, per = light ? 3 : 1, groups_per_batch = CH / step;
    int32_t* sa = (int32_t*)malloc(sizeof(int32_t) * groups_per_batch * 3);
    int32_t* sb = (int32_t*)malloc(sizeof(int32_t) * groups_per_batch * 3);
    uint8_t* img = NULL;
    size_t img_cap = 0;
    u32 i = start_frame;
    int done = 0;
    while (!done && !rc) {
        u32 base = i, g = 0, ne = 0;
        while (g < groups_per_batch) {
            u32 o = i - base;
            if (light) { sa[ne] = o; sb[ne] = -1; sa[ne + 1] = o + 1; sb[ne + 1] = o + 2; sa[ne + 2] = o + 3; sb[ne + 2] = -1; }
            else { sa[ne] = o; sb[ne] = o + 1; }
            ne += per; g++; i += step;
            if (i + 4 >= end_frame || i > end_frame) { done = 1; break; }
        }
        u32 nf = g * step; /* source frames base .. base+nf-1 (HEAVY loads but ignores the 3rd and 4th frame of a group) */
        for (u32 k = 0; k < nf && !rc; k++) {
            if (!light && base + k > end_frame) break;
            if (cur_dir) snprintf(path, sizeof path, "%s%lu.bmp", basename, base + k);
            else snprintf(path, sizeof path, "%s/%s%lu.bmp", dir, basename, base + k);
            if (load_bmp(path, width, height, buf + (size_t)k * SP)) rc = AGMVB_ERR_FILE;
        }
        uint64_t nbytes = 0;
        if (!rc) rc = agmvb_enc_frames(c, buf, nf, 0, sa, sb, ne, (uint32_t)encoded, &nbytes);
        if (!rc && nbytes > img_cap) { free(img); img_cap = nbytes * 2; img = (uint8_t*)malloc(img_cap); }
        if (!rc) rc = agmvb_enc_fetch(c, img, img_cap, NULL, NULL);
        if (!rc) fwrite(img, 1, nbytes, file);
        encoded += ne;
        agmv->frame_count += ne;
    }
    if (rc) fail(rc, "AGMV_EncodeAGMV (encode pass)");
    /* back-patch (:3615-3620) */
    fseek(file, 4, SEEK_SET);
    w32(file, encoded);
    fseek(file, 18, SEEK_SET);
    f32 rate = (f32)adjusted / (agmv->header.num_of_frames + 1);
    w32(file, (u32)round(agmv->header.frames_per_second * rate));
    fclose(file);
    free(buf); free(sa); free(sb); free(img);
    agmvb_enc_set_audio(c, NULL, 0, 16, 0, 0, 0);
    agmvb_enc_set_audio_stub(c, 1);
    DestroyAGMV(agmv); /* the reference consumes the caller's handle (:3625) */

    if (opt == AGMV_OPT_GBA_I || opt == AGMV_OPT_GBA_II || opt == AGMV_OPT_GBA_III) { /* :3627-3656 */
        FILE* in = fopen(filename, "rb");
        if (!in) return;
        fseek(in, 0, SEEK_END);
        long size = ftell(in);
        fseek(in, 0, SEEK_SET);
        uint8_t* data = (uint8_t*)malloc((size_t)size);
        if (fread(data, 1, (size_t)size, in) != (size_t)size) { }
        fclose(in);
        FILE* out = fopen("GBA_GEN_AGMV.h", "w");
        if (out) {
            fprintf(out, "#ifndef GBA_GEN_AGMV_H\n#define GBA_GEN_AGMV_H\n\nconst unsigned char GBA_AGMV_FILE[%ld] = {\n", size);
            for (long k = 0; k < size; k++) {
                if (k != 0 && k % 4000 == 0) fprintf(out, "\n");
                fprintf(out, "%d,", data[k]);
            }
            fprintf(out, "};\n\n#endif");
            fclose(out);
        }
        free(data);
    }
}

/* the GBA profiles also dump the stream as a C header (src/agmv_encode.c:3627-3656) */
static void export_gba_header(const char* filename) {
    FILE* in = fopen(filename, "rb");
    if (!in) return;
    fseek(in, 0, SEEK_END);
    long size = ftell(in);
    fseek(in, 0, SEEK_SET);
    uint8_t* data = (uint8_t*)malloc((size_t)size);
    if (fread(data, 1, (size_t)size, in) != (size_t)size) { }
    fclose(in);
    FILE* out = fopen("GBA_GEN_AGMV.h", "w");
    if (out) {
        fprintf(out, "#ifndef GBA_GEN_AGMV_H\n#define GBA_GEN_AGMV_H\n\nconst unsigned char GBA_AGMV_FILE[%ld] = {\n", size);
        for (long k = 0; k < size; k++) {
            if (k != 0 && k % 4000 == 0) fprintf(out, "\n");
            fprintf(out, "%d,", data[k]);
        }
        fprintf(out, "};\n\n#endif");
        fclose(out);
    }
    free(data);
}

/* AGMV_EncodeVideo (src/agmv_encode.c:719-2268) and AGMV_EncodeFullAGMV (:3659-4407), BMP input: the whole sequence is
 * loaded and handed to the GPU layer in one call (the similarity-gated schedule needs every consecutive pair). */
static void encode_whole(int video, AGMV* agmv, const char* filename, const char* dir, const char* basename, u8 img_type, u32 start_frame,
                         u32 end_frame, u32 width, u32 height, u32 fps, AGMV_OPT opt, AGMV_QUALITY quality, AGMV_COMPRESSION compression) {
    agmvb_ctx* c = ctx_get();
    if (!c) { if (agmv) DestroyAGMV(agmv); return; }
    if (img_type != AGMV_IMG_BMP) {
        fprintf(stderr, "libagmv_dropin: only BMP input is served (image decoding is AGIDL's, out of scope)\n");
        if (agmv) DestroyAGMV(agmv);
        return;
    }
    const u32 n_src = end_frame - start_frame + 1;
    const size_t SP = (size_t)width * height;
    const int cur_dir = dir[0] == 'c' && dir[1] == 'u' && dir[2] == 'r';
    uint32_t* buf = (uint32_t*)malloc((size_t)n_src * SP * 4);
    char path[512];
    int rc = 0;
    for (u32 k = 0; k < n_src && !rc; k++) {
        if (cur_dir) snprintf(path, sizeof path, "%s%lu.bmp", basename, start_frame + k);
        else snprintf(path, sizeof path, "%s/%s%lu.bmp", dir, basename, start_frame + k);
        if (load_bmp(path, width, height, buf + (size_t)k * SP)) rc = AGMVB_ERR_FILE;
    }
    const uint64_t track_bytes = (!video && agmv && agmv->header.total_audio_duration != 0) ? (uint64_t)agmv->header.audio_size : 0;
    if (!rc && track_bytes) rc = track_to_gpu(c, agmv); /* AGMV_EncodeFullAGMV interleaves the track only when it has a duration (:4022-4029) */
    else if (!rc) rc = agmvb_enc_set_audio(c, NULL, 0, 16, 0, 0, 0);
    uint64_t cap = 4096 + (uint64_t)n_src * (SP * 3 + 64 + 8) + track_bytes, len = 0;
    uint8_t* out = (uint8_t*)malloc(cap);
    uint32_t nenc = 0;
    if (!rc) {
        if (video) rc = agmvb_encode_video(c, buf, 0, (uint32_t)n_src, (uint32_t)width, (uint32_t)height, (uint32_t)fps, (int)opt, (int)quality,
                                           (int)compression, out, cap, &len, &nenc);
        else rc = agmvb_encode_full(c, buf, 0, (uint32_t)n_src, (uint32_t)width, (uint32_t)height, (uint32_t)agmv->header.num_of_frames,
                                    (uint32_t)agmv->header.frames_per_second, (int)opt, (int)quality, (int)compression, out, cap, &len, &nenc);
    }
    g_enc.valid = 0;
    if (rc) fail(rc, video ? "AGMV_EncodeVideo" : "AGMV_EncodeFullAGMV");
    else {
        FILE* f = fopen(filename, "wb");
        if (f) { fwrite(out, 1, (size_t)len, f); fclose(f); }
    }
    free(buf); free(out);
    if (agmv) DestroyAGMV(agmv);
    if (!rc && (opt == AGMV_OPT_GBA_I || opt == AGMV_OPT_GBA_II || opt == AGMV_OPT_GBA_III)) export_gba_header(filename);
}

void AGMV_EncodeVideo(const char* filename, const char* dir, const char* basename, u8 img_type, u32 start_frame, u32 end_frame, u32 width,
                      u32 height, u32 frames_per_second, AGMV_OPT opt, AGMV_QUALITY quality, AGMV_COMPRESSION compression) {
    encode_whole(1, NULL, filename, dir, basename, img_type, start_frame, end_frame, width, height, frames_per_second, opt, quality, compression);
}

void AGMV_EncodeFullAGMV(AGMV* agmv, const char* filename, const char* dir, const char* basename, u8 img_type, u32 start_frame, u32 end_frame,
                         u32 width, u32 height, u32 frames_per_second, AGMV_OPT opt, AGMV_QUALITY quality, AGMV_COMPRESSION compression) {
    encode_whole(0, agmv, filename, dir, basename, img_type, start_frame, end_frame, width, height, frames_per_second, opt, quality, compression);
}

/* ------------------------------------------------------------------------------------ */
/* decode                                                                                  */
/* ------------------------------------------------------------------------------------ */
/* src/agmv_decode.c:91-143 */
int AGMV_DecodeHeader(FILE* file, AGMV* agmv) {
    if (fread(agmv->header.fourcc, 1, 4, file) != 4) { }
    agmv->header.num_of_frames = r32(file);
    agmv->header.width = r32(file);
    agmv->header.height = r32(file);
    agmv->header.fmt = (u8)r8(file);
    agmv->header.version = (u8)r8(file);
    agmv->header.frames_per_second = r32(file);
    agmv->header.total_audio_duration = r32(file);
    agmv->header.sample_rate = r32(file);
    agmv->header.audio_size = r32(file);
    agmv->header.num_of_channels = (u16)r16(file);
    agmv->header.bits_per_sample = (u16)r16(file);
    u8 v = agmv->header.version;
    if (memcmp(agmv->header.fourcc, "AGMV", 4) || !(v >= 1 && v <= 4) || agmv->header.frames_per_second >= 200 ||
        !(agmv->header.bits_per_sample == 16 || agmv->header.bits_per_sample == 8))
        return INVALID_HEADER_FORMATTING_ERR;
    for (int p = 0; p < ((v == 1 || v == 3) ? 2 : 1); p++) {
        u32* pal = p ? agmv->header.palette1 : agmv->header.palette0;
        for (int i = 0; i < 256; i++) { u32 r = r8(file), g = r8(file), b = r8(file); pal[i] = r << 16 | g << 8 | b; }
    }
    return NO_ERR;
}

/* decoder state on the GPU for handles driven through AGMV_DecodeFrameChunk */
#define MAX_BOUND 16
static struct { AGMV* key; int stream; u32 w, h; int version; uint32_t palsum; } g_bound[MAX_BOUND];

static uint32_t pal_sum(const AGMV* a) {
    uint32_t s = 2166136261u;
    for (int i = 0; i < 256; i++) { s = (s ^ (uint32_t)a->header.palette0[i]) * 16777619u; s = (s ^ (uint32_t)a->header.palette1[i]) * 16777619u; }
    return s;
}

/* ---- frame-ahead queue (the loop it serves: AGMV_PlayAGMV, src/agmv_playback.c:102-115) ----
 * A caller that takes frame chunks one after the other (every player does) gets the next AGMV_B200_AHEAD chunks decoded in one
 * batch behind its back; the following calls are answered from host memory as long as they arrive where a sequential
 * caller arrives (same FILE*, the predicted 'AGFC' offset, frame_count one higher). Anything else - a seek, a changed
 * frame_count, another file - withdraws the frames not taken yet: the decoder state is put back to what it was before the
 * batch and the frames the caller did take are decoded again, so the handle is in exactly the state the reference's would be. */
#define PLAY_AHEAD_MAX 32
typedef struct { long pos; u32 frame_count, frame_num, usize, csize; uint32_t bpos, consumed; } PlayEnt;
typedef struct {
    FILE* file;
    int n, head;                 /* frames decoded in the current batch / already handed out */
    PlayEnt ent[PLAY_AHEAD_MAX];
    uint64_t poff[PLAY_AHEAD_MAX];
    uint32_t us[PLAY_AHEAD_MAX], cs[PLAY_AHEAD_MAX], bp[PLAY_AHEAD_MAX], co[PLAY_AHEAD_MAX];
    uint32_t* px; size_t px_cap;  /* n * P pixels */
    uint8_t* slab; size_t slab_len, slab_cap; long slab_pos;
    int streak;                  /* sequential calls in a row */
    long last_end; u32 last_fc;  /* cursor and frame_count the previous call left behind */
} PlayQ;
static PlayQ g_play[MAX_BOUND];
static unsigned long g_play_hits, g_play_batches, g_play_rollbacks;
/* test / tuning hook: frames answered from the queue, batches decoded ahead, batches withdrawn */
void AGMV_B200_PlayQueueStats(unsigned long* hits, unsigned long* batches, unsigned long* rollbacks) {
    if (hits) *hits = g_play_hits;
    if (batches) *batches = g_play_batches;
    if (rollbacks) *rollbacks = g_play_rollbacks;
}
static int play_ahead(void) {   /* read per call: a player may change its mind (and so do the tests) */
    const char* e = getenv("AGMV_B200_AHEAD");
    int v = e ? atoi(e) : 8;
    if (v < 1) v = 1;
    if (v > PLAY_AHEAD_MAX) v = PLAY_AHEAD_MAX;
    return v;
}
static void play_reset(PlayQ* q) { q->n = q->head = 0; q->streak = 0; q->file = NULL; }

static int bound_stream_slot(agmvb_ctx* c, AGMV* agmv, int* stream, int* slot_out, int* fresh) {
    int slot = -1;
    *fresh = 0;
    uint32_t ps = pal_sum(agmv);
    for (int k = 0; k < MAX_BOUND; k++) if (g_bound[k].key == agmv) slot = k;
    if (slot >= 0) {
        if (g_bound[slot].w == agmv->header.width && g_bound[slot].h == agmv->header.height &&
            g_bound[slot].version == agmv->header.version && g_bound[slot].palsum == ps && agmv->frame_count != 0) {
            *stream = g_bound[slot].stream;
            *slot_out = slot;
            return 0;
        }
        agmvb_dec_close(c, g_bound[slot].stream); /* new stream on an old handle (or a rewind to frame 0) */
        g_bound[slot].key = NULL;
    }
    for (int k = 0; k < MAX_BOUND && slot < 0; k++) if (!g_bound[k].key) slot = k;
    if (slot < 0) { agmvb_dec_close(c, g_bound[0].stream); slot = 0; }
    play_reset(&g_play[slot]);
    *fresh = 1;
    *slot_out = slot;
    uint32_t p0[256], p1[256];
    for (int i = 0; i < 256; i++) { p0[i] = (uint32_t)agmv->header.palette0[i]; p1[i] = (uint32_t)agmv->header.palette1[i]; }
    int rc = agmvb_dec_open_raw(c, (uint32_t)agmv->header.width, (uint32_t)agmv->header.height, agmv->header.version, p0, p1, stream);
    if (rc) return rc;
    g_bound[slot].key = agmv; g_bound[slot].stream = *stream; g_bound[slot].w = agmv->header.width; g_bound[slot].h = agmv->header.height;
    g_bound[slot].version = agmv->header.version; g_bound[slot].palsum = ps;
    return 0;
}

/* hand frame `e` of the queue to the caller: everything AGMV_DecodeFrameChunk leaves in the handle and the FILE */
static void play_serve(PlayQ* q, int k, FILE* file, AGMV* agmv) {
    const PlayEnt* e = &q->ent[k];
    const size_t P = (size_t)agmv->header.width * agmv->header.height;
    const uint32_t* px = q->px + (size_t)k * P;
    memcpy(agmv->frame_chunk->fourcc, "AGFC", 4);
    agmv->frame_chunk->frame_num = e->frame_num;
    agmv->frame_chunk->uncompressed_size = e->usize;
    agmv->frame_chunk->compressed_size = e->csize;
    for (size_t i = 0; i < P; i++) agmv->frame->img_data[i] = px[i];
    if (agmv->frame_count % 4 == 0) for (size_t i = 0; i < P; i++) agmv->iframe->img_data[i] = px[i];
    agmv->bitstream->pos = e->bpos;
    agmv->frame_count++;
    fseek(file, e->pos + 16 + (long)e->consumed, SEEK_SET); /* where the reference's bit reader leaves the cursor */
    q->last_end = e->pos + 16 + (long)e->consumed;
    q->last_fc = (u32)agmv->frame_count;
}

/* the caller left the predicted path with frames of the batch not taken: decoder state back to before the batch, then the
 * frames that were taken once more */
static int play_withdraw(agmvb_ctx* c, PlayQ* q, int stream) {
    int rc = 0;
    if (q->head < q->n) {
        g_play_rollbacks++;
        rc = agmvb_dec_restore(c, stream);
        if (!rc && q->head > 0)
            rc = agmvb_dec_chunks(c, stream, q->slab, q->slab_len, (uint32_t)q->head, q->poff, q->us, q->cs, q->ent[0].frame_count, q->px, NULL, NULL);
    }
    q->n = q->head = 0;
    q->streak = 0;
    return rc;
}

/* make q->slab hold the file from q->slab_pos up to (at least) relative offset `upto`; returns bytes available */
static size_t slab_fill(PlayQ* q, FILE* file, size_t upto) {
    if (upto <= q->slab_len) return q->slab_len;
    size_t want = upto + (256u << 10);
    if (want > q->slab_cap) {
        uint8_t* nb = (uint8_t*)realloc(q->slab, want * 2);
        if (!nb) return q->slab_len;
        q->slab = nb; q->slab_cap = want * 2;
    }
    if (fseek(file, q->slab_pos + (long)q->slab_len, SEEK_SET) != 0) return q->slab_len;
    q->slab_len += fread(q->slab + q->slab_len, 1, want - q->slab_len, file);
    return q->slab_len;
}

/* src/agmv_decode.c:145-410 */
int AGMV_DecodeFrameChunk(FILE* file, AGMV* agmv) {
    agmvb_ctx* c = ctx_get();
    const long pos0 = ftell(file);
    const size_t P = (size_t)agmv->header.width * agmv->header.height;
    int stream = -1, slot = 0, fresh = 0;
    int rcb = c ? bound_stream_slot(c, agmv, &stream, &slot, &fresh) : 0;
    PlayQ* q = &g_play[slot];
    if (c && !rcb && q->head < q->n) {
        const PlayEnt* e = &q->ent[q->head];
        if (q->file == file && e->pos == pos0 && e->frame_count == (u32)agmv->frame_count) { /* the sequential caller */
            agmv->bitstream->pos = 0;
            play_serve(q, q->head++, file, agmv);
            g_play_hits++;
            return NO_ERR;
        }
        rcb = play_withdraw(c, q, stream);
        fseek(file, pos0, SEEK_SET);
    }
    agmv->bitstream->pos = 0;
    read_fourcc(file, agmv->frame_chunk->fourcc);
    agmv->frame_chunk->frame_num = r32(file);
    agmv->frame_chunk->uncompressed_size = r32(file);
    agmv->frame_chunk->compressed_size = r32(file);
    if (memcmp(agmv->frame_chunk->fourcc, "AGFC", 4)) return INVALID_HEADER_FORMATTING_ERR;
    if (!c) return MEMORY_CORRUPTION_ERR;
    const u32 usize = agmv->frame_chunk->uncompressed_size, csize = agmv->frame_chunk->compressed_size;
    const long data_start = ftell(file);
    /* csize comes from the file: never allocate past what the file still holds */
    long here = data_start, end_pos = data_start;
    if (fseek(file, 0, SEEK_END) == 0) { end_pos = ftell(file); fseek(file, here, SEEK_SET); }
    size_t left = end_pos > here ? (size_t)(end_pos - here) : 0;
    /* a caller that keeps taking the next chunk gets the following ones decoded with this one */
    const int sequential = !fresh && q->streak >= 0 && (u32)agmv->frame_count == q->last_fc && pos0 >= q->last_end && q->last_fc != 0;
    q->streak = sequential ? q->streak + 1 : 0;
    const int ahead = play_ahead();
    if (!rcb && ahead > 1 && q->streak >= 2 && (size_t)csize <= left && (uint64_t)usize <= 2 * (uint64_t)P + 64) {
        /* collect up to `ahead` chunks: the next one is looked for where AGMV_FindNextFrameChunk (src/agmv_utils.c:140-166)
         * would find it for a caller whose cursor stands after this payload */
        q->slab_pos = pos0; q->slab_len = 0; q->file = file;
        int n = 0;
        size_t at = 0; /* slab offset of chunk n's 'AGFC' */
        while (n < ahead) {
            if (slab_fill(q, file, at + 16) < at + 16) break;
            const uint8_t* h = q->slab + at;
            if (memcmp(h, "AGFC", 4)) break;
            uint32_t fn, us, cs;
            memcpy(&fn, h + 4, 4); memcpy(&us, h + 8, 4); memcpy(&cs, h + 12, 4);
            if ((uint64_t)us > 2 * (uint64_t)P + 64) break;
            const size_t pay_end = at + 16 + (size_t)cs;
            const size_t have = slab_fill(q, file, pay_end + 64);
            if (have < pay_end && n > 0) break; /* a payload that runs past the end of the file is left to the single path */
            q->ent[n].pos = pos0 + (long)at; q->ent[n].frame_count = (u32)agmv->frame_count + (u32)n;
            q->ent[n].frame_num = fn; q->ent[n].usize = us; q->ent[n].csize = cs;
            q->poff[n] = at + 16; q->us[n] = us; q->cs[n] = cs;
            n++;
            /* next 'AGFC' at or after the end of this payload (look at most 1 MB further) */
            size_t scan = pay_end, found = (size_t)-1;
            while (found == (size_t)-1 && scan < pay_end + (1u << 20)) {
                const size_t got = slab_fill(q, file, scan + 4096);
                if (got < scan + 4) break;
                const size_t lim = got - 3 < scan + 4096 ? got - 3 : scan + 4096;
                for (size_t i = scan; i < lim; i++) if (q->slab[i] == 'A' && !memcmp(q->slab + i, "AGFC", 4)) { found = i; break; }
                scan = lim;
            }
            if (found == (size_t)-1) break;
            at = found;
        }
        if (n > 1) {
            if (q->px_cap < (size_t)n * P) {
                free(q->px);
                q->px = (uint32_t*)malloc((size_t)ahead * P * 4);
                q->px_cap = q->px ? (size_t)ahead * P : 0;
            }
            int rc = q->px ? agmvb_dec_snapshot(c, stream) : 1;
            if (!rc) {
                rc = agmvb_dec_chunks(c, stream, q->slab, q->slab_len, (uint32_t)n, q->poff, q->us, q->cs, (uint32_t)agmv->frame_count, q->px, q->bp, q->co);
                if (rc) agmvb_dec_restore(c, stream); /* a damaged chunk further on: this frame alone, below */
            }
            if (!rc) {
                for (int k = 0; k < n; k++) { q->ent[k].bpos = q->bp[k]; q->ent[k].consumed = q->co[k]; }
                q->n = n; q->head = 1;
                g_play_batches++;
                play_serve(q, 0, file, agmv);
                return NO_ERR;
            }
        }
        q->n = q->head = 0;
        fseek(file, data_start, SEEK_SET);
    }
    size_t want = (size_t)csize + 64;
    if (want > left + 64) want = left + 64;
    uint8_t* pay = (uint8_t*)malloc(want);
    uint32_t* px = (uint32_t*)malloc(P * 4);
    if (!pay || !px) { free(pay); free(px); return MEMORY_CORRUPTION_ERR; }
    size_t got = fread(pay, 1, want, file);
    int rc = rcb;
    uint32_t bpos = 0, consumed = 0;
    if (!rc) rc = agmvb_dec_chunk(c, stream, pay, got, (uint32_t)usize, (uint32_t)csize, (uint32_t)agmv->frame_count, px, &bpos, &consumed);
    if (!rc) {
        for (size_t i = 0; i < P; i++) agmv->frame->img_data[i] = px[i];
        if (agmv->frame_count % 4 == 0) for (size_t i = 0; i < P; i++) agmv->iframe->img_data[i] = px[i];
        agmv->bitstream->pos = bpos;
        agmv->frame_count++;
        fseek(file, data_start + (long)consumed, SEEK_SET); /* where the reference's bit reader leaves the cursor */
        q->last_end = data_start + (long)consumed;
        q->last_fc = (u32)agmv->frame_count;
    }
    free(pay); free(px);
    if (rc) {
        fseek(file, data_start, SEEK_SET); /* nothing was consumed */
        fail(rc, "AGMV_DecodeFrameChunk");
        return rc <= 3 ? rc : MEMORY_CORRUPTION_ERR;
    }
    return NO_ERR;
}

/* src/agmv_decode.c:412-453: one 'AGAC' chunk, expanded on the GPU into audio_track->pcm / pcm8 at start_point */
int AGMV_DecodeAudioChunk(FILE* file, AGMV* agmv) {
    read_fourcc(file, agmv->audio_chunk->fourcc);
    agmv->audio_chunk->size = r32(file);
    if (memcmp(agmv->audio_chunk->fourcc, "AGAC", 4)) return INVALID_HEADER_FORMATTING_ERR;
    const size_t size = (size_t)agmv->audio_chunk->size;
    if (size == 0) return NO_ERR;
    agmvb_ctx* c = ctx_get();
    if (!c) return MEMORY_CORRUPTION_ERR;
    uint8_t* b = (uint8_t*)malloc(size);
    if (!b) return MEMORY_CORRUPTION_ERR; /* a damaged size field */
    size_t got = fread(b, 1, size, file);
    if (got < size) memset(b + got, 0xFF, size - got); /* AGIDL_ReadByte at the end of the file: EOF stored in a u8 */
    const int bits = agmv->header.bits_per_sample == 16 ? 16 : 8;
    void* dst = bits == 16 ? (void*)(agmv->audio_track->pcm + agmv->audio_track->start_point) : (void*)(agmv->audio_track->pcm8 + agmv->audio_track->start_point);
    int rc = agmvb_audio_expand(c, b, (uint64_t)size, bits, dst);
    free(b);
    if (rc) { fail(rc, "AGMV_DecodeAudioChunk"); return MEMORY_CORRUPTION_ERR; }
    agmv->audio_track->start_point += size;
    return NO_ERR;
}

/* src/agmv_utils.c:1403-1435, the WAV case (weak: the reference's agmv_utils.o brings AIFF / AIFC as well) */
WEAK void AGMV_ExportAudioType(FILE* audio, AGMV* agmv, AGMV_AUDIO_TYPE audio_type) {
    if (audio_type != AGMV_AUDIO_WAV) fprintf(stderr, "libagmv_dropin: AIFF export is agmv_utils.c's (link the reference's agmv_utils.o); writing WAV\n");
    const u32 bytes = agmv->header.bits_per_sample == 16 ? agmv->header.audio_size * 2 : agmv->header.audio_size;
    fwrite("RIFF", 1, 4, audio);
    w32(audio, bytes);
    fwrite("WAVEfmt ", 1, 8, audio);
    w32(audio, 16);
    w16(audio, 1);
    w16(audio, agmv->header.num_of_channels);
    w32(audio, agmv->header.sample_rate);
    w32(audio, 75600);
    w16(audio, (agmv->header.num_of_channels * agmv->header.bits_per_sample) / 8);
    w16(audio, agmv->header.bits_per_sample);
    fwrite("data", 1, 4, audio);
    w32(audio, bytes);
    if (agmv->header.bits_per_sample == 16) fwrite(agmv->audio_track->pcm, 2, (size_t)agmv->header.audio_size, audio);
    else fwrite(agmv->audio_track->pcm8, 1, (size_t)agmv->header.audio_size, audio);
}

/* src/agmv_decode.c:572-618: the whole track in one GPU call, then quick_export.wav / .aiff. Samples the stream's chunks
 * do not cover (audio_size - size * frames of them) are zero here; the reference leaves them uninitialised. */
static int export_track(agmvb_ctx* c, int stream, const uint8_t* hdr, AGMV_AUDIO_TYPE audio_type) {
    AGMV a;
    AGMV_AUDIO_TRACK tr;
    memset(&a, 0, sizeof a.header);
    a.header.total_audio_duration = hdr[22] | hdr[23] << 8 | hdr[24] << 16 | (u32)hdr[25] << 24;
    a.header.sample_rate = hdr[26] | hdr[27] << 8 | hdr[28] << 16 | (u32)hdr[29] << 24;
    a.header.audio_size = hdr[30] | hdr[31] << 8 | hdr[32] << 16 | (u32)hdr[33] << 24;
    a.header.num_of_channels = (u16)(hdr[34] | hdr[35] << 8);
    a.header.bits_per_sample = (u16)(hdr[36] | hdr[37] << 8);
    if (a.header.total_audio_duration == 0) return 0;
    uint64_t n = 0;
    int bits = 16;
    int rc = agmvb_dec_audio(c, stream, NULL, 0, &n, &bits);
    if (rc) return rc;
    const uint64_t cap = n > a.header.audio_size ? n : a.header.audio_size;
    void* pcm = calloc((size_t)cap + 1, 2);
    if (!pcm) return MEMORY_CORRUPTION_ERR; /* a damaged audio_size field */
    rc = agmvb_dec_audio(c, stream, pcm, cap, &n, &bits);
    if (!rc) {
        tr.pcm = (u16*)pcm; tr.pcm8 = (u8*)pcm; tr.start_point = n; tr.total_audio_duration = 0;
        a.audio_track = &tr;
        const int aiff = audio_type == AGMV_AUDIO_AIFF || audio_type == AGMV_AUDIO_AIFC;
        FILE* f = fopen(aiff ? "quick_export.aiff" : "quick_export.wav", "wb");
        if (f) { AGMV_ExportAudioType(f, &a, aiff ? audio_type : AGMV_AUDIO_WAV); fclose(f); }
    }
    free(pcm);
    return rc;
}

/* src/agmv_decode.c:527-647 (AGMV_DecodeAGMV), :455-525 (AGMV_DecodeVideo: frames only), :682-767 (AGMV_DecodeAudio: track only) */
static int decode_file(const char* filename, u8 img_type, AGMV_AUDIO_TYPE audio_type, int frames, int audio, const char* who) {
    FILE* f = fopen(filename, "rb");
    if (!f) return FILE_NOT_FOUND_ERR;
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* data = (uint8_t*)malloc((size_t)size + 1);
    if (!data) { fclose(f); return MEMORY_CORRUPTION_ERR; }
    if (fread(data, 1, (size_t)size, f) != (size_t)size) { }
    fclose(f);
    agmvb_ctx* c = ctx_get();
    if (!c) { free(data); return MEMORY_CORRUPTION_ERR; }
    int stream = -1;
    uint32_t w = 0, h = 0, n = 0;
    int rc = agmvb_dec_open(c, data, (uint64_t)size, &stream, &w, &h, &n);
    uint8_t hdr[38];
    memcpy(hdr, data, size >= 38 ? 38 : 0);
    free(data);
    if (rc) { fail(rc, who); return rc <= 3 ? rc : MEMORY_CORRUPTION_ERR; }
    if (frames) {
        const size_t P = (size_t)w * h;
        const uint32_t CH = 32;
        uint32_t* px = (uint32_t*)malloc((size_t)CH * P * 4);
        if (!px) rc = AGMVB_ERR_MEMORY;
        for (uint32_t f0 = 0; f0 < n && !rc; f0 += CH) {
            uint32_t nf = n - f0 < CH ? n - f0 : CH;
            rc = agmvb_dec_frames(c, stream, nf, px, 0);
            if (!rc && img_type == AGMV_IMG_BMP) for (uint32_t k = 0; k < nf; k++) quick_export_bmp(px + (size_t)k * P, w, h);
        }
        free(px);
    }
    if (!rc && audio) rc = export_track(c, stream, hdr, audio_type);
    agmvb_dec_close(c, stream);
    if (rc) { fail(rc, who); return rc <= 3 ? rc : MEMORY_CORRUPTION_ERR; }
    return NO_ERR;
}
int AGMV_DecodeAGMV(const char* filename, u8 img_type, AGMV_AUDIO_TYPE audio_type) { return decode_file(filename, img_type, audio_type, 1, 1, "AGMV_DecodeAGMV"); }
int AGMV_DecodeVideo(const char* filename, u8 img_type) { return decode_file(filename, img_type, AGMV_AUDIO_WAV, 1, 0, "AGMV_DecodeVideo"); }
int AGMV_DecodeAudio(const char* filename, AGMV_AUDIO_TYPE audio_type) { return decode_file(filename, 0, audio_type, 0, 1, "AGMV_DecodeAudio"); }

/* src/agmv_decode.c:649-680: an unsigned integer as the 80-bit IEEE extended float of an AIFF header (sample rate);
 * the reference's agmv_utils.c (AIFF / AIFC export, :1457,1510) calls it, so a drop-in for agmv_decode.c has to export it.
 * Sign 0, exponent 16383 + floor(log2 num), mantissa = num with its top bit at bit 63; only bytes 0..5 are written. */
void to_80bitfloat(u32 num, u8 bytes[10]) {
    if (num <= 1) { bytes[0] = 0x3F; bytes[1] = 0xFF; bytes[2] = 0x80; return; }
    bytes[0] = 0x40;
    if (num >= 0x40000000ul) { bytes[1] = 0x1D; return; }
    int lead = 0;   /* zero bits below bit 31 before the number's highest set bit, counted from bit 30 down */
    while (!(num & (0x40000000ul >> lead))) lead++;
    const unsigned long mant = lead < 31 ? num << (lead + 1) : 0;
    bytes[1] = (u8)(29 - lead);
    bytes[2] = (u8)(mant >> 24);
    bytes[3] = (u8)(mant >> 16);
    bytes[4] = (u8)(mant >> 8);
    bytes[5] = (u8)mant;
}
