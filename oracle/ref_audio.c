/* TEST INFRASTRUCTURE ONLY - driver around the UNMODIFIED reference audio chunk codec.
 *
 *   ref_audio compress IN OUT     IN = n little-endian u16 samples; runs AGMV_CompressAudio (src/agmv_encode.c:659-705)
 *                                 on a handle from CreateAGMV with bits_per_sample = 16; OUT = n bytes (atsample).
 *   ref_audio expand IN OUT       IN = n atsample bytes; wraps them in one 'AGAC' chunk and runs AGMV_DecodeAudioChunk
 *                                 (src/agmv_decode.c:412-453) with bits_per_sample = 16; OUT = n little-endian u16.
 *   ref_audio track FILE OUT      what AGMV_DecodeAGMV does for a stream with audio (src/agmv_decode.c:572-600) without the
 *                                 picture export: per frame AGMV_FindNextFrameChunk, AGMV_DecodeFrameChunk,
 *                                 AGMV_FindNextAudioChunk, AGMV_DecodeAudioChunk; OUT = the start_point samples written
 *                                 (u16 for 16-bit tracks, u8 otherwise). Prints "rc R samples S bits B".
 * Links against oracle/_ref/libagmv_ref.so (compiled from /root/reference where it lies).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <agmv.h>

static long fsize(FILE* f) { fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); return n; }

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s compress|expand|track IN OUT\n", argv[0]); return 2; }
    FILE* in = fopen(argv[2], "rb");
    if (!in) return 2;
    if (!strcmp(argv[1], "compress")) {
        long n = fsize(in) / 2;
        AGMV* agmv = CreateAGMV(1, 4, 4, 1);
        AGMV_SetBitsPerSample(agmv, 16);
        AGMV_SetAudioSize(agmv, n);
        agmv->audio_track->pcm = (u16*)malloc(2 * n + 2);
        if (fread(agmv->audio_track->pcm, 2, n, in) != (size_t)n) return 2;
        agmv->audio_chunk->atsample = (u8*)malloc(n + 1);
        AGMV_CompressAudio(agmv);
        FILE* out = fopen(argv[3], "wb");
        fwrite(agmv->audio_chunk->atsample, 1, n, out);
        fclose(out);
        return 0;
    }
    if (!strcmp(argv[1], "expand")) {
        long n = fsize(in);
        u8* b = (u8*)malloc(n + 1);
        if (fread(b, 1, n, in) != (size_t)n) return 2;
        FILE* tmp = tmpfile();
        fwrite("AGAC", 1, 4, tmp);
        AGMV_WriteLong(tmp, n);
        fwrite(b, 1, n, tmp);
        fseek(tmp, 0, SEEK_SET);
        AGMV* agmv = CreateAGMV(1, 4, 4, 1);
        AGMV_SetBitsPerSample(agmv, 16);
        agmv->audio_track->pcm = (u16*)calloc(n + 1, 2);
        agmv->audio_track->start_point = 0;
        int rc = AGMV_DecodeAudioChunk(tmp, agmv);
        FILE* out = fopen(argv[3], "wb");
        fwrite(agmv->audio_track->pcm, 2, n, out);
        fclose(out);
        printf("rc %d\n", rc);
        return 0;
    }
    /* track */
    AGMV* agmv = (AGMV*)calloc(1, sizeof(AGMV));
    agmv->frame_chunk = (AGMV_FRAME_CHUNK*)calloc(1, sizeof(AGMV_FRAME_CHUNK));
    agmv->audio_chunk = (AGMV_AUDIO_CHUNK*)calloc(1, sizeof(AGMV_AUDIO_CHUNK));
    agmv->bitstream = (AGMV_BITSTREAM*)calloc(1, sizeof(AGMV_BITSTREAM));
    agmv->frame = (AGMV_FRAME*)calloc(1, sizeof(AGMV_FRAME));
    agmv->iframe = (AGMV_FRAME*)calloc(1, sizeof(AGMV_FRAME));
    agmv->audio_track = (AGMV_AUDIO_TRACK*)calloc(1, sizeof(AGMV_AUDIO_TRACK));
    int err = AGMV_DecodeHeader(in, agmv);
    if (err != NO_ERR) { printf("rc %d\n", err); return 0; }
    size_t px = (size_t)agmv->header.width * agmv->header.height;
    agmv->frame->width = agmv->iframe->width = agmv->header.width;
    agmv->frame->height = agmv->iframe->height = agmv->header.height;
    agmv->frame->img_data = (u32*)calloc(px, sizeof(u32));
    agmv->iframe->img_data = (u32*)calloc(px, sizeof(u32));
    agmv->bitstream->len = px * 2;
    agmv->bitstream->data = (u8*)calloc(px * 2, 1);
    int bits = agmv->header.bits_per_sample;
    agmv->audio_track->pcm = (u16*)calloc(agmv->header.audio_size + 1, 2);
    agmv->audio_track->pcm8 = (u8*)calloc(agmv->header.audio_size + 1, 1);
    agmv->audio_track->start_point = 0;
    unsigned long n = AGMV_GetNumberOfFrames(agmv), i;
    for (i = 0; i < n && err == NO_ERR; i++) {
        AGMV_FindNextFrameChunk(in);
        err = AGMV_DecodeFrameChunk(in, agmv);
        if (err != NO_ERR) break;
        AGMV_FindNextAudioChunk(in);
        err = AGMV_DecodeAudioChunk(in, agmv);
    }
    FILE* out = fopen(argv[3], "wb");
    if (bits == 16) fwrite(agmv->audio_track->pcm, 2, agmv->audio_track->start_point, out);
    else fwrite(agmv->audio_track->pcm8, 1, agmv->audio_track->start_point, out);
    fclose(out);
    printf("rc %d samples %lu bits %d\n", err, agmv->audio_track->start_point, bits);
    return 0;
}
