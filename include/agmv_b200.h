/* agmv_b200.h - C-ABI of libagmv_b200.so: the B200 (sm_100a) implementation of
 * libagmv's per-frame encode / decode hot path.
 *
 * Plain C: opaque context, raw pointers and sizes, int status codes; no CUDA or
 * torch types in any signature (a CUDA stream crosses as void*). Pixels are
 * 32-bit 0x00RRGGBB words; the drop-in layer (agmv_dropin.h) narrows the
 * reference's 8-byte `u32` pixels before they get here.
 *
 * The library has NO CPU fallback: every entry point that computes launches
 * CUDA kernels and fails with AGMVB_ERR_CUDA when no device is usable.
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to the reference checkout).
 */
#ifndef AGMV_B200_H
#define AGMV_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes: 0..3 are the reference's `enum Error` (include/agmv_defines.h:37-42) */
#define AGMVB_OK 0
#define AGMVB_ERR_HEADER 1      /* INVALID_HEADER_FORMATTING_ERR */
#define AGMVB_ERR_FILE 2        /* FILE_NOT_FOUND_ERR */
#define AGMVB_ERR_MEMORY 3      /* MEMORY_CORRUPTION_ERR */
#define AGMVB_ERR_ARG 4
#define AGMVB_ERR_CUDA 5
#define AGMVB_ERR_UNSUPPORTED 6

/* AGMV_OPT / AGMV_QUALITY / AGMV_COMPRESSION values, include/agmv_defines.h:56-76 */
enum { AGMVB_OPT_I = 1, AGMVB_OPT_II, AGMVB_OPT_III, AGMVB_OPT_ANIM, AGMVB_OPT_GBA_I, AGMVB_OPT_GBA_II, AGMVB_OPT_GBA_III, AGMVB_OPT_NDS };
enum { AGMVB_HIGH_QUALITY = 1, AGMVB_MID_QUALITY, AGMVB_LOW_QUALITY };
enum { AGMVB_LZSS = 1, AGMVB_LZ77 };

typedef struct agmvb_ctx agmvb_ctx;

/* ---- context ------------------------------------------------------------- */
/* device: CUDA ordinal. stream: a cudaStream_t to launch on (NULL = the context
 * creates its own). Replaces CreateAGMV's allocations (src/agmv_utils.c:332-369)
 * for the device side. */
int agmvb_create(agmvb_ctx** out, int device, void* cuda_stream);
void agmvb_destroy(agmvb_ctx* ctx);
const char* agmvb_last_error(const agmvb_ctx* ctx);
/* kernels launched by this context so far (bench.py's gpu_launches) */
uint64_t agmvb_kernel_launches(const agmvb_ctx* ctx);
/* wait for everything queued on the context's stream */
int agmvb_sync(agmvb_ctx* ctx);
/* Pixel format of HOST frame buffers (on_device == 0): every `frames` argument of the encoder entry points and the host
 * `out` of agmvb_dec_frames. AGMVB_PIX_U32 (default): one uint32 0x00RRGGBB per pixel. AGMVB_PIX_BGR24: packed B,G,R bytes,
 * row-major, no padding - exactly the pixel rows of the 24-bit BMP files the reference reads (AGIDL_LoadBMP) and writes
 * (AGIDL_QuickExport); a quarter fewer bytes across PCIe, unpacked / packed on the device. Device buffers are always u32. */
#define AGMVB_PIX_U32 0
#define AGMVB_PIX_BGR24 1
int agmvb_set_host_format(agmvb_ctx* ctx, int fmt);

/* ---- encoder: the pieces of AGMV_EncodeAGMV (src/agmv_encode.c:2270-3657) --- */
/* Start a sequence: source size, profile, quality, entropy coder. GBA / NDS
 * profiles code at 120x80 / 128x96 (include/agmv_encode.h:21-24). Clears the
 * histogram, the colour memo table and the frame counter. */
int agmvb_enc_begin(agmvb_ctx* ctx, uint32_t src_w, uint32_t src_h, int opt, int quality, int compression);

/* Pass 1 (:2371-2568): add n_frames source frames to the quantised-colour
 * histogram. `frames` is host memory unless on_device != 0. */
int agmvb_enc_histogram(agmvb_ctx* ctx, const uint32_t* frames, uint64_t n_frames, int on_device);
/* Device pointer / bin count of the 64-bit histogram, for the caller's
 * all-reduce when the sequence is sharded over several GPUs. */
int agmvb_enc_histogram_ptr(agmvb_ctx* ctx, uint64_t** dev_bins, uint32_t* n_bins);

/* AGMV_BubbleSort + greedy pick + split (src/agmv_utils.c:995-1010,
 * src/agmv_encode.c:2570-2656) on the device. */
int agmvb_enc_build_palette(agmvb_ctx* ctx);
int agmvb_enc_get_palette(agmvb_ctx* ctx, uint32_t pal0[256], uint32_t pal1[256]);
/* AGMV_SetICP0 / AGMV_SetICP1 (src/agmv_utils.c:259-271) for callers that
 * bring their own palettes (the per-frame API). */
int agmvb_enc_set_palette(agmvb_ctx* ctx, const uint32_t pal0[256], const uint32_t pal1[256]);

/* I-frame entry state of the per-frame API (agmv->iframe_entries,
 * src/agmv_encode.c:626-630): (pal_num << 8 | index) per coded pixel. */
int agmvb_enc_set_iframe_entries(agmvb_ctx* ctx, const uint16_t* entries);
int agmvb_enc_get_iframe_entries(agmvb_ctx* ctx, uint16_t* entries);

/* Pass 2: encode n_enc frames = AGMV_EncodeFrame (+ the empty AGMV_EncodeAudioChunk)
 * n_enc times (src/agmv_encode.c:529-634, :707-717). Encoded frame k is
 * frames[src_a[k]] when src_b[k] < 0, else AGMV_InterpFrame(frames[src_a[k]],
 * frames[src_b[k]]) (src/agmv_utils.c:949-969). first_frame_count is
 * agmv->frame_count of the first one (I-frame iff frame_count % 4 == 0; the
 * AGFC header carries frame_count + 1). The result - for every frame
 * 'AGFC' hdr | csize payload bytes | 8 x 0xFF | 'AGAC' 0 - stays on the device;
 * *image_bytes receives its length. */
int agmvb_enc_frames(agmvb_ctx* ctx, const uint32_t* frames, uint64_t n_frames_in_buffer, int on_device,
                     const int32_t* src_a, const int32_t* src_b, uint32_t n_enc, uint32_t first_frame_count,
                     uint64_t* image_bytes);
/* Copy the last agmvb_enc_frames result to host: chunk image and per-frame
 * usize / csize (either size array may be NULL). */
int agmvb_enc_fetch(agmvb_ctx* ctx, uint8_t* image, uint64_t cap, uint32_t* usize, uint32_t* csize);
int agmvb_enc_image_ptr(agmvb_ctx* ctx, uint8_t** dev_image, uint64_t* bytes);

/* The whole of AGMV_EncodeAGMV on in-memory source frames start..end
 * (n_src = end - start + 1): histogram, palette, header (src/agmv_encode.c:21-94),
 * PDIFS schedule (:2727-2770, :3610-3612), frame chunks, header back-patch
 * (:3615-3620). Writes the .agmv image to `out` (host). */
int agmvb_encode_sequence(agmvb_ctx* ctx, const uint32_t* frames, int on_device, uint32_t n_src, uint32_t w, uint32_t h,
                          uint32_t create_n, uint32_t fps, int opt, int quality, int compression,
                          uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* n_encoded);
/* agmvb_encode_sequence over several GPUs of this process: ctxs[k] = one context per device (agmvb_create with
 * that device); frames in HOST memory. Every GPU histograms its share of the source frames, the bins are summed over peer
 * copies (the palette is global: src/agmv_encode.c:2371-2568), every GPU encodes a GOP-aligned range of the PDIFS schedule
 * (:2727-2770) and copies its chunks to their final place in `out`. The bytes are those of the single-GPU call.
 * AGMVB_LZSS only (the LZ77 coder carries state from frame to frame, :218-224). */
int agmvb_encode_sequence_multi(agmvb_ctx* const* ctxs, int n_ctx, const uint32_t* frames, uint32_t n_src, uint32_t w, uint32_t h,
                                uint32_t create_n, uint32_t fps, int opt, int quality, int compression,
                                uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* n_encoded);
/* Range of encoded frames [first, first + count) that shard `shard` of `n_shards` takes: borders are multiples of 12 encoded
 * frames for the LIGHT profiles (4 PDIFS groups = 3 GOPs = 16 source frames), of 4 for the HEAVY ones. */
int agmvb_shard_range(uint32_t n_enc, int light, int n_shards, int shard, uint32_t* first, uint32_t* count);
/* The reference's other two sequence encoders, same conventions as agmvb_encode_sequence:
 * AGMV_EncodeVideo (src/agmv_encode.c:719-2268): a pair of frames is merged only if the fraction of grey-equal pixels
 * (AGMV_CompareFrameSimilarity, src/agmv_utils.c:920-947) reaches the profile's leniency; no audio chunks;
 * AGMV_EncodeFullAGMV (src/agmv_encode.c:3659-4407): every frame encoded, no audio chunks, header not back-patched. */
int agmvb_encode_video(agmvb_ctx* ctx, const uint32_t* frames, int on_device, uint32_t n_src, uint32_t w, uint32_t h, uint32_t fps,
                       int opt, int quality, int compression, uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* n_encoded);
int agmvb_encode_full(agmvb_ctx* ctx, const uint32_t* frames, int on_device, uint32_t n_src, uint32_t w, uint32_t h,
                      uint32_t create_n, uint32_t fps, int opt, int quality, int compression,
                      uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t* n_encoded);
/* AGMV_CompareFrameSimilarity's numerator for n_pairs frame pairs (indices into `frames`, coded size). */
int agmvb_frame_similarity(agmvb_ctx* ctx, const uint32_t* frames, uint64_t n_frames_in_buffer, int on_device,
                           const int32_t* pair_a, const int32_t* pair_b, uint32_t n_pairs, uint64_t* counts);
/* Whether agmvb_enc_frames appends the empty 'AGAC' chunk after every frame chunk (AGMV_EncodeAGMV does, the per-frame
 * AGMV_EncodeFrame and the other two encoders do not). Default on. */
int agmvb_enc_set_audio_stub(agmvb_ctx* ctx, int on);
/* ---- audio chunk codec -------------------------------------------------------- */
/* The audio track of the handle about to be encoded, as AGMV_WavToAudioTrack leaves it (src/agmv_utils.c:1035-1087):
 * audio_size samples of 16-bit (uint16_t) or 8-bit PCM in host memory plus the header fields. Uploads the track and runs
 * AGMV_CompressAudio (src/agmv_encode.c:659-705: every 16-bit sample becomes one byte - the closest of a square, a rounded
 * square and a shifted code; 8-bit samples are copied) on the device, where the result stays. The next
 * agmvb_encode_sequence / agmvb_encode_full writes the header's audio fields and one 'AGAC' chunk of
 * (u32)(audio_size / (f32)frames) bytes after every frame chunk (AGMV_EncodeAudioChunk, :707-717; frames = the adjusted
 * frame count of :2296-2353 for AGMV_EncodeAGMV, end - start for AGMV_EncodeFullAGMV) and consumes the track.
 * pcm == NULL or audio_size == 0 clears it. The track survives agmvb_enc_begin. */
int agmvb_enc_set_audio(agmvb_ctx* ctx, const void* pcm, uint64_t audio_size, int bits_per_sample, uint32_t sample_rate,
                        uint32_t channels, uint32_t total_duration);
/* For callers that drive agmvb_enc_frames themselves: from now on frame number g (first_frame_count + k) is followed by
 * 'AGAC' chunk_size and atsample[g * chunk_size .. +chunk_size) (bytes past the end of the track read as 0; the reference
 * reads past its allocation there). agmvb_enc_set_audio_stub switches back to empty / no audio chunks. */
int agmvb_enc_set_audio_chunk(agmvb_ctx* ctx, uint32_t chunk_size);
/* the companded track (audio_size bytes), e.g. for AGMV_EncodeAudioChunk calls of the per-frame API */
int agmvb_enc_get_atsample(agmvb_ctx* ctx, uint8_t* out, uint64_t cap);
/* AGMV_CompressAudio / the sample loop of AGMV_DecodeAudioChunk (src/agmv_decode.c:431-451: even byte b -> b*b, odd byte
 * -> b << 8, i.e. AGMV_SQR_TABLE / AGMV_SHIFT_TABLE) on n samples, host buffers. */
int agmvb_audio_compress(agmvb_ctx* ctx, const void* pcm, uint64_t n, int bits_per_sample, uint8_t* atsample);
int agmvb_audio_expand(agmvb_ctx* ctx, const uint8_t* atsample, uint64_t n, int bits_per_sample, void* pcm);

/* Header only (AGMV_EncodeHeader, src/agmv_encode.c:21-94) with the current palette. */
int agmvb_enc_header(agmvb_ctx* ctx, uint32_t n_frames, uint32_t fps, uint8_t* out, uint64_t cap, uint64_t* len);

/* ---- decoder: AGMV_DecodeAGMV / AGMV_DecodeFrameChunk (src/agmv_decode.c:91-410, 527-647) --- */
/* Parse header + palettes (AGMV_DecodeHeader), index the frame chunks the way
 * AGMV_FindNextFrameChunk walks them (src/agmv_utils.c:140-166), upload the
 * stream. Returns a stream handle (>= 0) in *stream. Several streams may be
 * open at once (batched decode of independent streams). */
int agmvb_dec_open(agmvb_ctx* ctx, const uint8_t* file, uint64_t len, int* stream,
                   uint32_t* w, uint32_t* h, uint32_t* n_frames);
/* Decode the next `count` frames of one stream into out (count*w*h pixels;
 * host memory unless on_device). Frames must be taken in order: the decoder
 * state (pixels, I-frame, expanded-bitstream leftovers, frame_count) carries
 * over exactly like the reference's AGMV handle. */
int agmvb_dec_frames(agmvb_ctx* ctx, int stream, uint32_t count, uint32_t* out, int on_device);
/* Batched decode: the next `count` frames of each of n_streams streams of
 * equal size. outs[s] is a device pointer (count*w*h pixels) or NULL to keep
 * only a ring of recent frames on the device; checksums (nullable, host,
 * n_streams*count) receives a 64-bit position-weighted sum per frame. */
int agmvb_dec_batch(agmvb_ctx* ctx, const int* streams, uint32_t n_streams, uint32_t count,
                    uint32_t* const* outs, uint64_t* checksums);
/* The audio half of AGMV_DecodeAGMV (src/agmv_decode.c:572-587): every 'AGAC' chunk of an open stream, decoded by
 * AGMV_DecodeAudioChunk's rule into pcm (uint16_t samples for 16-bit tracks, bytes for 8-bit ones; host). *n_samples =
 * audio_track->start_point after the last chunk. pcm == NULL only queries the sizes. Independent of the frame cursor. */
int agmvb_dec_audio(agmvb_ctx* ctx, int stream, void* pcm, uint64_t cap_samples, uint64_t* n_samples, int* bits_per_sample);
int agmvb_dec_close(agmvb_ctx* ctx, int stream);
/* Seek: what AGMV_SkipTo / AGMV_SkipBackwards do to the handle (src/agmv_playback.c:81-100) - the next
 * frame decoded is frame_index, with frame_count = frame_index; the chunk offsets come from the index built by
 * agmvb_dec_open (the reference's AGMV_ParseAGMV offset_table, src/agmv_utils.c:219-243). Decoder state (pixels,
 * I-frame snapshot, expanded-bitstream leftovers) is NOT rewound, as in the reference: seek to an I-frame
 * (frame_index % 4 == 0) for a clean picture. */
int agmvb_dec_seek(agmvb_ctx* ctx, int stream, uint32_t frame_index);
/* Streaming entry: one AGMV_DecodeFrameChunk (src/agmv_decode.c:145-410) at a time, for callers that own the
 * FILE* (AGMV_PlayAGMV, src/agmv_playback.c:102-115). agmvb_dec_open_raw takes what AGMV_DecodeHeader left in
 * the handle (size, version, palettes); agmvb_dec_chunk takes the bytes after the 16-byte 'AGFC' header
 * (csize payload bytes plus at least the 8-byte trailer), agmv->frame_count, and returns the frame (host,
 * w*h pixels), bitstream->pos and the number of payload bytes the bit reader consumed (the new file cursor). */
int agmvb_dec_open_raw(agmvb_ctx* ctx, uint32_t w, uint32_t h, int version, const uint32_t pal0[256], const uint32_t pal1[256], int* stream);
int agmvb_dec_chunk(agmvb_ctx* ctx, int stream, const uint8_t* payload, uint64_t payload_len, uint32_t usize, uint32_t csize,
                    uint32_t frame_count, uint32_t* out_px, uint32_t* bpos, uint32_t* consumed);
/* The frame-ahead queue behind AGMV_PlayAGMV's loop (src/agmv_playback.c:102-115): n consecutive frame
 * chunks in one call. slab = a piece of the file that holds all n chunks; payload_off[k] = slab offset of the first byte
 * after chunk k's 16-byte 'AGFC' header; frames are decoded with frame_count = first_frame_count + k. Outputs per frame:
 * pixels (host, n*w*h), bitstream->pos and payload bytes consumed (bpos / consumed: n entries each, nullable). */
int agmvb_dec_chunks(agmvb_ctx* ctx, int stream, const uint8_t* slab, uint64_t slab_len, uint32_t n, const uint64_t* payload_off,
                     const uint32_t* usize, const uint32_t* csize, uint32_t first_frame_count, uint32_t* out_px, uint32_t* bpos,
                     uint32_t* consumed);
/* Put the decoder state of a stream aside / bring it back (pixels, I-frame snapshot, carried bitstream buffer - what the
 * reference keeps in the AGMV handle, include/agmv_defines.h:131-162), so that frames decoded ahead of the caller can be
 * withdrawn when the caller seeks (AGMV_SkipTo, src/agmv_playback.c:94-100) instead of taking them. */
int agmvb_dec_snapshot(agmvb_ctx* ctx, int stream);
int agmvb_dec_restore(agmvb_ctx* ctx, int stream);

/* ---- measurement utilities ------------------------------------------------------ */
/* Fill dev_out with n synthetic w x h frames t = first_t .. first_t+n-1 (the deterministic integer generator
 * of SURVEY.md 8d; same formula as oracle/agmv_oracle.c). Benchmark input only. */
int agmvb_synth_frames(agmvb_ctx* ctx, uint32_t* dev_out, uint32_t w, uint32_t h, uint32_t first_t, uint32_t n, uint32_t seed);
/* Bracket every kernel launch with CUDA events on the launching stream and read back, per kernel class,
 * the launch count and the summed device milliseconds (arrays of agmvb_profile_classes() entries). */
int agmvb_profile(agmvb_ctx* ctx, int enable);
int agmvb_profile_classes(void);
const char* agmvb_profile_name(int cls);
int agmvb_profile_read(agmvb_ctx* ctx, uint64_t* counts, double* total_ms);

/* ---- unit-test hooks (thin wrappers over single kernels) ---------------------- */
/* white-box access to the last decoded chunk's internal buffers (0 bpos, 1 block records, 2 consumed, 3 stale bytes, 4 expansions) */
int agmvb_test_peek(agmvb_ctx* ctx, int which, void* dst, uint64_t bytes);
/* AGMV_LZSS (src/agmv_encode.c:106-177) over F byte buffers concatenated in
 * `data` (frame_start has F+1 entries). out receives, per buffer, the bytes the
 * reference writes (ceil(outbits/8), at out + out_off[f]); csize/outbits as the
 * reference computes them. */
int agmvb_test_lzss(agmvb_ctx* ctx, const uint8_t* data, const uint32_t* frame_start, uint32_t F,
                    uint8_t* out, uint64_t out_cap, uint64_t* out_off, uint32_t* csize, uint32_t* outbits);
/* AGMV_LZ77 (src/agmv_encode.c:179-238) over F buffers that share one carried bitstream buffer, as consecutive frames of
 * one handle do; every byte of that buffer starts as persist_fill (the reference: 0). out + out_off[f]: frame f's tokens. */
int agmvb_test_lz77(agmvb_ctx* ctx, const uint8_t* data, const uint32_t* frame_start, uint32_t F, int persist_fill,
                    uint8_t* out, uint64_t out_cap, uint64_t* out_off, uint32_t* csize);
/* AGMV_FindNearestEntry / AGMV_FindNearestColor over n colours (src/agmv_utils.c:785-895). dual: bit 0 = two palettes;
 * bit 1 = compute the entries of all 2^24 colours first, as calls that encode long sequences do, and read the answers there */
int agmvb_test_quantize(agmvb_ctx* ctx, const uint32_t* colors, uint64_t n, const uint32_t pal0[256],
                        const uint32_t pal1[256], int dual, uint16_t* entries);
/* AGMV_Assemble{I,P}FrameBitstream on ready-made entries (src/agmv_encode.c:354-527) */
int agmvb_test_assemble(agmvb_ctx* ctx, const uint16_t* entries, const uint16_t* iframe_entries /* NULL: I-frame */,
                        uint32_t w, uint32_t h, int dual, const uint32_t pal0[256], const uint32_t pal1[256],
                        uint8_t* out, uint64_t cap, uint32_t* usize);

#ifdef __cplusplus
}
#endif
#endif
