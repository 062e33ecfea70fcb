/* TEST INFRASTRUCTURE ONLY - driver around the UNMODIFIED reference encoder.
 *
 * One AGMV_EncodeAGMV call per fresh process (the
 * reference's output depends on fresh zeroed mmap pages, so it must not share
 * a process with anything else). Links against oracle/_ref/libagmv_ref.so,
 * which is compiled from /root/reference where it lies.
 *
 *   ref_encode OUT DIR BASE START END W H FPS OPT QUALITY COMPRESSION CREATE_N [MODE]
 *   MODE: agmv (default) = AGMV_EncodeAGMV, video = AGMV_EncodeVideo, full = AGMV_EncodeFullAGMV
 *   environment REF_WAV=file.wav: AGMV_WavToAudioTrack(file.wav, agmv) between CreateAGMV and the encode call, as in
 *   examples/simple_video_and_audio/simple_video_and_audio.c:20-22 (agmv and full modes)
 *
 * mirrors the call sequence of the reference's own examples
 * (examples/simple_video/simple_video.c: CreateAGMV then AGMV_EncodeAGMV).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <agmv.h>

int main(int argc, char** argv) {
    if (argc < 13) {
        fprintf(stderr, "usage: %s out dir base start end w h fps opt quality compression create_n\n", argv[0]);
        return 2;
    }
    const char* out = argv[1];
    const char* dir = argv[2];
    const char* base = argv[3];
    unsigned long start = strtoul(argv[4], 0, 10), end = strtoul(argv[5], 0, 10);
    unsigned long w = strtoul(argv[6], 0, 10), h = strtoul(argv[7], 0, 10), fps = strtoul(argv[8], 0, 10);
    int opt = atoi(argv[9]), quality = atoi(argv[10]), comp = atoi(argv[11]);
    unsigned long create_n = strtoul(argv[12], 0, 10);

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const char* mode = argc > 13 ? argv[13] : "agmv";
    if (!strcmp(mode, "video")) {
        AGMV_EncodeVideo(out, dir, base, AGMV_IMG_BMP, start, end, w, h, fps, (AGMV_OPT)opt, (AGMV_QUALITY)quality, (AGMV_COMPRESSION)comp);
    } else {
        AGMV* agmv = CreateAGMV(create_n, w, h, fps);
        if (getenv("REF_WAV")) AGMV_WavToAudioTrack(getenv("REF_WAV"), agmv);
        if (!strcmp(mode, "full"))
            AGMV_EncodeFullAGMV(agmv, out, dir, base, AGMV_IMG_BMP, start, end, w, h, fps, (AGMV_OPT)opt, (AGMV_QUALITY)quality, (AGMV_COMPRESSION)comp);
        else
            AGMV_EncodeAGMV(agmv, out, dir, base, AGMV_IMG_BMP, start, end, w, h, fps, (AGMV_OPT)opt, (AGMV_QUALITY)quality, (AGMV_COMPRESSION)comp);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    fprintf(stderr, "ref_encode_seconds %.6f\n",
            (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
    return 0;
}
