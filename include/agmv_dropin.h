/* agmv_dropin.h - the reference's own C API for the frame hot path, served by the
 * B200 implementation (libagmv_dropin.so -> libagmv_b200.so).
 *
 * libagmv has no plugin layer: its public API *is* the boundary. A
 * replacement therefore has to be binary compatible with translation units compiled
 * against the reference's headers. This header restates exactly the part of that
 * interface the hot path touches - the fundamental types, the handle layout and the
 * entry points - so that
 *   (a) code written against the reference compiles unchanged against this file, and
 *   (b) objects compiled against the reference's include/agmv_defines.h can be linked
 *       with libagmv_dropin.so in place of agmv_encode.o / agmv_decode.o
 *       (INTEGRATION.md shows the link line).
 * Layout facts that matter (LP64): u32 is `unsigned long` = 8 bytes, so a pixel handed
 * across this API is 8 bytes wide, sizeof(AGMV_ENTRY) == 16, sizeof(AGMV) == 324264.
 * tests/test_dropin.py checks those numbers against the reference build.
 *
 * Each declaration cites the reference declaration it mirrors.
 */
#ifndef AGMV_DROPIN_H
#define AGMV_DROPIN_H
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* fundamental types: include/agmv_defines.h:20-30 */
typedef unsigned char u8;
typedef unsigned short u16;
typedef unsigned long u32;
typedef signed char s8;
typedef signed short s16;
typedef signed long s32;
typedef float f32;
typedef int Bool;

/* include/agmv_defines.h:37-42 */
typedef enum Error { NO_ERR = 0, INVALID_HEADER_FORMATTING_ERR = 1, FILE_NOT_FOUND_ERR = 2, MEMORY_CORRUPTION_ERR = 3 } Error;

#define MAX_OFFSET_TABLE 40000 /* include/agmv_defines.h:47 */

/* include/agmv_defines.h:56-76 */
typedef enum AGMV_OPT {
    AGMV_OPT_I = 1, AGMV_OPT_II = 2, AGMV_OPT_III = 3, AGMV_OPT_ANIM = 4,
    AGMV_OPT_GBA_I = 5, AGMV_OPT_GBA_II = 6, AGMV_OPT_GBA_III = 7, AGMV_OPT_NDS = 8
} AGMV_OPT;
typedef enum AGMV_QUALITY { AGMV_HIGH_QUALITY = 1, AGMV_MID_QUALITY = 2, AGMV_LOW_QUALITY = 3 } AGMV_QUALITY;
typedef enum AGMV_COMPRESSION { AGMV_LZSS_COMPRESSION = 1, AGMV_LZ77_COMPRESSION = 2 } AGMV_COMPRESSION;

/* include/agmv_defines.h:78-93 */
typedef struct AGMV_MAIN_HEADER {
    char fourcc[4];
    u32 num_of_frames, width, height;
    u8 fmt, version;
    u32 frames_per_second, total_audio_duration, sample_rate, audio_size;
    u16 num_of_channels, bits_per_sample;
    u32 palette0[256];
    u32 palette1[256];
} AGMV_MAIN_HEADER;

/* include/agmv_defines.h:95-100 */
typedef struct AGMV_FRAME_CHUNK { char fourcc[4]; u32 frame_num, uncompressed_size, compressed_size; } AGMV_FRAME_CHUNK;
/* include/agmv_defines.h:102-107 */
typedef struct AGMV_AUDIO_CHUNK { char fourcc[4]; u32 size; u8* atsample; s8* satsample; } AGMV_AUDIO_CHUNK;
/* include/agmv_defines.h:109-113 */
typedef struct AGMV_FRAME { u32 width, height; u32* img_data; } AGMV_FRAME;
/* include/agmv_defines.h:115-120 */
typedef struct AGMV_AUDIO_TRACK { u32 total_audio_duration, start_point; u16* pcm; u8* pcm8; } AGMV_AUDIO_TRACK;
/* include/agmv_defines.h:122-126 */
typedef struct AGMV_ENTRY { u8 pal_num, index; u32 occurence; } AGMV_ENTRY;
/* include/agmv_defines.h:140-144 */
typedef struct AGMV_BITSTREAM { u8* data; u32 len, pos; } AGMV_BITSTREAM;

/* the handle: include/agmv_defines.h:146-162 */
typedef struct AGMV {
    AGMV_MAIN_HEADER header;
    AGMV_FRAME_CHUNK* frame_chunk;
    AGMV_AUDIO_CHUNK* audio_chunk;
    AGMV_BITSTREAM* bitstream;
    AGMV_FRAME* frame;
    AGMV_FRAME* iframe;
    AGMV_AUDIO_TRACK* audio_track;
    AGMV_ENTRY* iframe_entries;
    AGMV_OPT opt;
    AGMV_COMPRESSION compression;
    u32 frame_count;
    f32 leniency;
    u32 offset_table[MAX_OFFSET_TABLE];
    Bool enable_audio;
    f32 volume;
} AGMV;

/* include/agmv_defines.h:164-183 (only BMP frames are served by this build) */
typedef enum AGMV_IMG_TYPE { AGMV_IMG_BMP = 1, AGMV_IMG_TGA = 2, AGMV_IMG_TIM = 3, AGMV_IMG_PCX = 4, AGMV_IMG_LMP = 5, AGMV_IMG_PVR = 6,
                             AGMV_IMG_GXT = 7, AGMV_IMG_BTI = 8, AGMV_IMG_3DF = 9, AGMV_IMG_PPM = 10, AGMV_IMG_LBM = 11 } AGMV_IMG_TYPE;
typedef enum AGMV_AUDIO_TYPE { AGMV_AUDIO_WAV = 1, AGMV_AUDIO_AIFF = 2, AGMV_AUDIO_AIFC = 3, AGMV_AUDIO_RAW = 4 } AGMV_AUDIO_TYPE;

/* handle lifecycle: include/agmv_utils.h:70-71, src/agmv_utils.c:332-421.
 * Defined WEAK here so that the reference's own agmv_utils.o wins when both are linked. */
AGMV* CreateAGMV(u32 num_of_frames, u32 width, u32 height, u32 frames_per_second);
void DestroyAGMV(AGMV* agmv);

/* encode: include/agmv_encode.h:28-29,36 */
void AGMV_EncodeHeader(FILE* file, AGMV* agmv);
void AGMV_EncodeFrame(FILE* file, AGMV* agmv, u32* img_data);
void AGMV_EncodeAGMV(AGMV* agmv, const char* filename, const char* dir, const char* basename, u8 img_type, u32 start_frame,
                     u32 end_frame, u32 width, u32 height, u32 frames_per_second, AGMV_OPT opt, AGMV_QUALITY quality,
                     AGMV_COMPRESSION compression);

/* the other two sequence encoders: include/agmv_encode.h:35,37 */
void AGMV_EncodeVideo(const char* filename, const char* dir, const char* basename, u8 img_type, u32 start_frame, u32 end_frame, u32 width,
                      u32 height, u32 frames_per_second, AGMV_OPT opt, AGMV_QUALITY quality, AGMV_COMPRESSION compression);
void AGMV_EncodeFullAGMV(AGMV* agmv, const char* filename, const char* dir, const char* basename, u8 img_type, u32 start_frame,
                         u32 end_frame, u32 width, u32 height, u32 frames_per_second, AGMV_OPT opt, AGMV_QUALITY quality,
                         AGMV_COMPRESSION compression);

/* audio chunk codec: include/agmv_encode.h (AGMV_CompressAudio, AGMV_EncodeAudioChunk), include/agmv_decode.h
 * (AGMV_DecodeAudioChunk), include/agmv_utils.h (AGMV_ExportAudioType; WAV only here, weak). The sample maps run on the GPU. */
void AGMV_CompressAudio(AGMV* agmv);
void AGMV_EncodeAudioChunk(FILE* file, AGMV* agmv);
int AGMV_DecodeAudioChunk(FILE* file, AGMV* agmv);
void AGMV_ExportAudioType(FILE* audio, AGMV* agmv, AGMV_AUDIO_TYPE audio_type);

/* decode: include/agmv_decode.h:21-22,26 */
int AGMV_DecodeHeader(FILE* file, AGMV* agmv);
int AGMV_DecodeFrameChunk(FILE* file, AGMV* agmv);
int AGMV_DecodeAGMV(const char* filename, u8 img_type, AGMV_AUDIO_TYPE audio_type);
/* the rest of include/agmv_decode.h:23-25, defined in src/agmv_decode.c:455-525 (frames only), :682-767 (audio track only) and
 * :649-680 (AIFF sample-rate field; called from the reference's agmv_utils.c:1457,1510, so whoever replaces agmv_decode.o owes it) */
int AGMV_DecodeVideo(const char* filename, u8 img_type);
int AGMV_DecodeAudio(const char* filename, AGMV_AUDIO_TYPE audio_type);
void to_80bitfloat(u32 num, u8 bytes[10]);

/* not part of the reference: last error text of the GPU layer, and the device to use (default 0) */
const char* AGMV_B200_LastError(void);
/* Frame-ahead queue behind AGMV_DecodeFrameChunk (what AGMV_PlayAGMV's loop, src/agmv_playback.c:102-115, calls per frame):
 * frames answered from the queue, batches decoded ahead, batches withdrawn because the caller went elsewhere.
 * Depth: environment AGMV_B200_AHEAD (default 8, 1 = off). */
void AGMV_B200_PlayQueueStats(unsigned long* hits, unsigned long* batches, unsigned long* rollbacks);
void AGMV_B200_SetDevice(int device);

#ifdef __cplusplus
}
#endif
#endif
