/* TEST INFRASTRUCTURE ONLY - CPU restatement of libagmv's frame hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this. The product (libagmv_b200.so) never
 * links or calls it.
 *
 * Parity status: PINNED. agmv_oracle.c is checked byte-for-byte against the
 * unmodified reference built by oracle/Makefile (oracle/_ref) and against the
 * committed golden vectors under tests/golden/ (tests/test_oracle.py).
 *
 * Pixels are 32-bit 0x00RRGGBB words (the reference's `u32` is 8 bytes on
 * LP64; only the low 24 bits ever carry data).
 */
#ifndef AGMV_ORACLE_H
#define AGMV_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/agmv_defines.h:56-76 of the reference */
enum { ORC_OPT_I = 1, ORC_OPT_II, ORC_OPT_III, ORC_OPT_ANIM, ORC_OPT_GBA_I, ORC_OPT_GBA_II, ORC_OPT_GBA_III, ORC_OPT_NDS };
enum { ORC_HIGH_QUALITY = 1, ORC_MID_QUALITY, ORC_LOW_QUALITY };
enum { ORC_LZSS = 1, ORC_LZ77 };

/* synthetic input generator (SURVEY.md 8d / BASELINE.md 4) */
void orc_synth_frame(int w, int h, int t, uint32_t seed, uint32_t* out);
int orc_write_bmp24(const char* path, int w, int h, const uint32_t* px);
int orc_read_bmp24(const char* path, int* w, int* h, uint32_t* px, size_t cap_px);

/* colour helpers */
uint32_t orc_max_clr(int quality);
uint32_t orc_quantize_color(uint32_t c, int quality);
int orc_opt_is_dual(int opt);
int orc_opt_is_light(int opt);

/* E0: histogram over raw counts (bins start at 0 here, at 1 in the reference:
 * the constant offset does not change the sort). hist has max_clr bins. */
void orc_histogram_add(uint64_t* hist, const uint32_t* px, size_t n, int quality);
/* E1-E3 */
void orc_build_palette(const uint64_t* hist, int quality, int opt, uint32_t pal0[256], uint32_t pal1[256]);
/* E4 */
void orc_interp(uint32_t* dst, const uint32_t* a, const uint32_t* b, size_t n);
/* E13 */
void orc_scale_nearest(const uint32_t* src, int w, int h, float sx, float sy, uint32_t* dst, int* nw, int* nh);
/* E5/E6: entries are (pal_num << 8) | index */
uint16_t orc_nearest_entry(const uint32_t* pal0, const uint32_t* pal1, int dual, uint32_t c);
void orc_quantize_frame(const uint32_t* px, size_t n, const uint32_t* pal0, const uint32_t* pal1, int dual, uint16_t* entries);
/* E7-E9: returns usize */
size_t orc_assemble(const uint16_t* ent, const uint16_t* ient, int w, int h, int is_iframe, int dual,
                    const uint32_t* pal0, const uint32_t* pal1, uint8_t* out);
/* E10: returns csize = (u32)((float)outbits/8.0f); writes ceil(outbits/8) bytes */
uint32_t orc_lzss(const uint8_t* in, size_t n, uint8_t* out, size_t* nbytes, uint64_t* outbits);
uint32_t orc_lz77(const uint8_t* in, size_t n, uint8_t* out, size_t* nbytes);

/* E8/E11/E12: whole AGMV_EncodeAGMV on in-memory source frames start..end
 * (n_src = end-start+1 frames of w*h pixels). Returns file length or <0. */
long orc_encode_agmv(const uint32_t* frames, int n_src, int w, int h, uint32_t create_n, uint32_t fps,
                     int opt, int quality, int compression, uint8_t* out, size_t out_cap);

/* "next" row N1: AGMV_EncodeVideo (similarity-gated PDIFS, src/agmv_encode.c:719-2268) and AGMV_EncodeFullAGMV (:3659-4407) */
long orc_encode_video(const uint32_t* frames, int n_src, int w, int h, uint32_t fps, int opt, int quality, int compression,
                      uint8_t* out, size_t out_cap);
long orc_encode_full(const uint32_t* frames, int n_src, int w, int h, uint32_t create_n, uint32_t fps,
                     int opt, int quality, int compression, uint8_t* out, size_t out_cap);

/* E11 for a range of encoded frames with a given palette (multi-GPU sharding tests) */
long orc_encode_frames(const uint32_t* frames, int w, int h, const int32_t* src_a, const int32_t* src_b, int n_enc, uint32_t first_fc,
                       const uint32_t* pal0, const uint32_t* pal1, int dual, uint8_t* out, size_t cap);

/* D1-D5: decode a whole .agmv image. frames_out receives n*w*h pixels (may be
 * NULL to query). Returns the reference's error code (0 ok). */
int orc_decode_agmv(const uint8_t* file, size_t len, uint32_t* frames_out, size_t cap_px,
                    int* w, int* h, int* n_frames);

/* "next" row N4: the audio chunk codec. orc_audio_compress16 = AGMV_CompressAudio's 16-bit branch (src/agmv_encode.c:659-705),
 * orc_audio_expand16 = the sample loop of AGMV_DecodeAudioChunk (src/agmv_decode.c:431-445). Pinned against the unmodified
 * reference on all 65 536 sample values and all 256 codes (tests/golden/audio_*.bin). */
void orc_audio_compress16(const uint16_t* pcm, size_t n, uint8_t* atsample);
void orc_audio_expand16(const uint8_t* atsample, size_t n, uint16_t* pcm);

#ifdef __cplusplus
}
#endif
#endif
