/* TEST INFRASTRUCTURE ONLY - driver around the UNMODIFIED reference decoder.
 *
 * One decode per fresh process. Two modes:
 *
 *   ref_decode export FILE        AGMV_DecodeAGMV(FILE, BMP, WAV): writes
 *                                 ./quick_export_<k>.bmp exactly as the
 *                                 reference does (src/agmv_decode.c:527-647);
 *                                 prints "rc <code>".
 *   ref_decode raw FILE OUT       per-frame public API without the BMP export
 *                                 (what AGMV_PlayAGMV does, src/agmv_playback.c:102-115):
 *                                 handle allocated as AGMV_DecodeAGMV does but
 *                                 zero-filled (calloc: frame_count := 0, buffers := 0),
 *                                 AGMV_DecodeHeader, then AGMV_FindNextFrameChunk +
 *                                 AGMV_DecodeFrameChunk per frame. Each frame is
 *                                 appended to OUT as W*H little-endian 32-bit
 *                                 0x00RRGGBB words (OUT "-" = no dump, timing only).
 *                                 Audio chunks are skipped by the chunk scan.
 *   ref_decode seek FILE OUT PLAN same handle, but the frames are visited in the order given by PLAN
 *                                 (comma-separated frame indices): AGMV_ParseAGMV fills agmv->offset_table
 *                                 (src/agmv_utils.c:219-243), and whenever the next index is not the
 *                                 successor of the previous one the driver does what AGMV_SkipTo does
 *                                 (src/agmv_playback.c:94-100, without rounding to an I-frame):
 *                                 fseek(offset_table[k]) and frame_count = k. Pixel, I-frame and bitstream
 *                                 buffers keep whatever the last decoded frame left in them.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include <agmv.h>

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s export FILE | raw FILE OUT\n", argv[0]);
        return 2;
    }
    if (!strcmp(argv[1], "export")) {
        double t0 = now();
        int rc = AGMV_DecodeAGMV(argv[2], AGMV_IMG_BMP, AGMV_AUDIO_WAV);
        fprintf(stderr, "ref_decode_seconds %.6f\n", now() - t0);
        printf("rc %d\n", rc);
        return 0;
    }
    if (argc < 4) return 2;

    AGMV* agmv = (AGMV*)calloc(1, sizeof(AGMV));
    agmv->frame_chunk = (AGMV_FRAME_CHUNK*)calloc(1, sizeof(AGMV_FRAME_CHUNK));
    agmv->audio_chunk = (AGMV_AUDIO_CHUNK*)calloc(1, sizeof(AGMV_AUDIO_CHUNK));
    agmv->bitstream = (AGMV_BITSTREAM*)calloc(1, sizeof(AGMV_BITSTREAM));
    agmv->frame = (AGMV_FRAME*)calloc(1, sizeof(AGMV_FRAME));
    agmv->iframe = (AGMV_FRAME*)calloc(1, sizeof(AGMV_FRAME));
    agmv->audio_track = (AGMV_AUDIO_TRACK*)calloc(1, sizeof(AGMV_AUDIO_TRACK));

    FILE* f = fopen(argv[2], "rb");
    if (!f) { printf("rc %d\n", FILE_NOT_FOUND_ERR); return 0; }
    int err = AGMV_DecodeHeader(f, agmv);
    if (err != NO_ERR) { printf("rc %d\n", err); return 0; }

    size_t px = (size_t)agmv->header.width * agmv->header.height;
    agmv->frame->width = agmv->iframe->width = agmv->header.width;
    agmv->frame->height = agmv->iframe->height = agmv->header.height;
    agmv->frame->img_data = (u32*)calloc(px, sizeof(u32));
    agmv->iframe->img_data = (u32*)calloc(px, sizeof(u32));
    agmv->bitstream->len = px * 2;
    agmv->bitstream->data = (u8*)calloc(px * 2, 1);

    FILE* out = strcmp(argv[3], "-") ? fopen(argv[3], "wb") : NULL;
    uint32_t* narrow = (uint32_t*)malloc(px * 4);
    unsigned long n = AGMV_GetNumberOfFrames(agmv), i;
    double dec = 0.0;
    if (!strcmp(argv[1], "seek")) {
        if (argc < 5) return 2;
        AGMV_ParseAGMV(f, agmv);
        long prev = -2;
        unsigned long done = 0;
        char* tok = strtok(argv[4], ",");
        while (tok) {
            long k = atol(tok);
            if (k < 0 || (unsigned long)k >= n) { err = MEMORY_CORRUPTION_ERR; break; }
            if (k != prev + 1) {
                fseek(f, agmv->offset_table[k], SEEK_SET);
                agmv->frame_count = k;
            }
            AGMV_FindNextFrameChunk(f);
            err = AGMV_DecodeFrameChunk(f, agmv);
            if (err != NO_ERR) break;
            if (out) {
                size_t q;
                for (q = 0; q < px; q++) narrow[q] = (uint32_t)agmv->frame->img_data[q];
                fwrite(narrow, 4, px, out);
            }
            prev = k;
            done++;
            tok = strtok(NULL, ",");
        }
        if (out) fclose(out);
        fclose(f);
        printf("rc %d frames %lu w %lu h %lu\n", err, done, agmv->header.width, agmv->header.height);
        return 0;
    }
    for (i = 0; i < n; i++) {
        double t0 = now();
        AGMV_FindNextFrameChunk(f);
        err = AGMV_DecodeFrameChunk(f, agmv);
        dec += now() - t0;
        if (err != NO_ERR) break;
        if (out) {
            size_t k;
            for (k = 0; k < px; k++) narrow[k] = (uint32_t)agmv->frame->img_data[k];
            fwrite(narrow, 4, px, out);
        }
    }
    if (out) fclose(out);
    fclose(f);
    fprintf(stderr, "ref_decode_seconds %.6f\n", dec);
    printf("rc %d frames %lu w %lu h %lu\n", err, i, agmv->header.width, agmv->header.height);
    return 0;
}
