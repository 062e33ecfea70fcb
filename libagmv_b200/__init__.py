"""libagmv_b200 - B200 (sm_100a) implementation of libagmv's frame hot path.

This package is a thin ctypes binding of the C-ABI in include/agmv_b200.h
(libagmv_b200.so: hand-written CUDA kernels + host orchestration in C++) and of
the reference-facing drop-in API (libagmv_dropin.so, include/agmv_dropin.h).
There is no Python or CPU implementation of the path here: if the shared
library is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libagmv_b200.so")
DROPIN_PATH = os.path.join(_HERE, "libagmv_dropin.so")

OPT = dict(I=1, II=2, III=3, ANIM=4, GBA_I=5, GBA_II=6, GBA_III=7, NDS=8)
QUALITY = dict(HIGH=1, MID=2, LOW=3)
LZSS, LZ77 = 1, 2

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)

# every symbol include/agmv_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "agmvb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "agmvb_destroy": (None, [C.c_void_p]),
    "agmvb_last_error": (C.c_char_p, [C.c_void_p]),
    "agmvb_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "agmvb_sync": (C.c_int, [C.c_void_p]),
    "agmvb_set_host_format": (C.c_int, [C.c_void_p, C.c_int]),
    "agmvb_enc_begin": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int]),
    "agmvb_enc_histogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "agmvb_enc_histogram_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), _u32p]),
    "agmvb_enc_build_palette": (C.c_int, [C.c_void_p]),
    "agmvb_enc_get_palette": (C.c_int, [C.c_void_p, _u32p, _u32p]),
    "agmvb_enc_set_palette": (C.c_int, [C.c_void_p, _u32p, _u32p]),
    "agmvb_enc_set_iframe_entries": (C.c_int, [C.c_void_p, _u16p]),
    "agmvb_enc_get_iframe_entries": (C.c_int, [C.c_void_p, _u16p]),
    "agmvb_enc_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, _i32p, _i32p, C.c_uint32, C.c_uint32, _u64p]),
    "agmvb_enc_fetch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, _u32p, _u32p]),
    "agmvb_enc_image_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), _u64p]),
    "agmvb_enc_header": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_uint64, _u64p]),
    "agmvb_encode_sequence": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, _u64p, _u32p]),
    "agmvb_encode_sequence_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, _u64p, _u32p]),
    "agmvb_shard_range": (C.c_int, [C.c_uint32, C.c_int, C.c_int, C.c_int, _u32p, _u32p]),
    "agmvb_encode_video": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_uint64, _u64p, _u32p]),
    "agmvb_encode_full": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_uint64, _u64p, _u32p]),
    "agmvb_frame_similarity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, _i32p, _i32p, C.c_uint32, _u64p]),
    "agmvb_enc_set_audio_stub": (C.c_int, [C.c_void_p, C.c_int]),
    "agmvb_enc_set_audio": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]),
    "agmvb_enc_set_audio_chunk": (C.c_int, [C.c_void_p, C.c_uint32]),
    "agmvb_enc_get_atsample": (C.c_int, [C.c_void_p, _u8p, C.c_uint64]),
    "agmvb_audio_compress": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, _u8p]),
    "agmvb_audio_expand": (C.c_int, [C.c_void_p, _u8p, C.c_uint64, C.c_int, C.c_void_p]),
    "agmvb_dec_audio": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, _u64p, C.POINTER(C.c_int)]),
    "agmvb_dec_open": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_int), _u32p, _u32p, _u32p]),
    "agmvb_dec_frames": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_int]),
    "agmvb_dec_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), _u64p]),
    "agmvb_dec_close": (C.c_int, [C.c_void_p, C.c_int]),
    "agmvb_dec_seek": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32]),
    "agmvb_dec_open_raw": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, _u32p, _u32p, C.POINTER(C.c_int)]),
    "agmvb_dec_chunk": (C.c_int, [C.c_void_p, C.c_int, _u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _u32p, _u32p, _u32p]),
    "agmvb_dec_chunks": (C.c_int, [C.c_void_p, C.c_int, _u8p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), _u32p, _u32p, C.c_uint32, _u32p, _u32p, _u32p]),
    "agmvb_dec_snapshot": (C.c_int, [C.c_void_p, C.c_int]),
    "agmvb_dec_restore": (C.c_int, [C.c_void_p, C.c_int]),
    "agmvb_synth_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "agmvb_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "agmvb_profile_classes": (C.c_int, []),
    "agmvb_profile_name": (C.c_char_p, [C.c_int]),
    "agmvb_profile_read": (C.c_int, [C.c_void_p, _u64p, C.POINTER(C.c_double)]),
    "agmvb_test_peek": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]),
    "agmvb_test_lzss": (C.c_int, [C.c_void_p, _u8p, _u32p, C.c_uint32, _u8p, C.c_uint64, _u64p, _u32p, _u32p]),
    "agmvb_test_lz77": (C.c_int, [C.c_void_p, _u8p, _u32p, C.c_uint32, C.c_int, _u8p, C.c_uint64, _u64p, _u32p]),
    "agmvb_test_quantize": (C.c_int, [C.c_void_p, _u32p, C.c_uint64, _u32p, _u32p, C.c_int, _u16p]),
    "agmvb_test_assemble": (C.c_int, [C.c_void_p, _u16p, _u16p, C.c_uint32, C.c_uint32, C.c_int, _u32p, _u32p, _u8p,
                                      C.c_uint64, _u32p]),
}

_lib = None


def load():
    """dlopen libagmv_b200.so and bind every declared symbol. Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class AgmvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"agmv_b200 error {code}: {msg}")
        self.code = code


def _p(a, t):
    return a.ctypes.data_as(t)


class Context:
    """One GPU context (device workspaces + a stream). `stream` is a raw cudaStream_t handle (int) or None."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.agmvb_create(C.byref(h), device, C.c_void_p(stream) if stream else None)
        if rc != 0:
            raise AgmvError(rc, "agmvb_create failed (no usable CUDA device?)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.agmvb_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise AgmvError(rc, self.lib.agmvb_last_error(self.h).decode())

    @property
    def launches(self):
        return int(self.lib.agmvb_kernel_launches(self.h))

    def sync(self):
        self._ck(self.lib.agmvb_sync(self.h))

    def set_host_format(self, fmt):
        """0 = uint32 0x00RRGGBB host pixels (default), 1 = packed B,G,R bytes (BMP pixel rows)."""
        self._ck(self.lib.agmvb_set_host_format(self.h, fmt))
        self.host_fmt = fmt

    # ---- encoder ----------------------------------------------------------
    def encode_sequence(self, frames, create_n, fps, opt, quality, compression=LZSS, device_ptr=None, shape=None, out=None):
        """AGMV_EncodeAGMV on in-memory frames. frames: (n,h,w) uint32 host array, or pass device_ptr + shape."""
        if device_ptr is None and getattr(self, "host_fmt", 0) == 1:
            frames = np.ascontiguousarray(frames, dtype=np.uint8)   # (n, h, w, 3) packed B,G,R
            n, h, w, _ = frames.shape
            src, on_dev = frames.ctypes.data, 0
        elif device_ptr is None:
            frames = np.ascontiguousarray(frames, dtype=np.uint32)
            n, h, w = frames.shape
            src, on_dev = frames.ctypes.data, 0
        else:
            n, h, w = shape
            src, on_dev = device_ptr, 1
        cap = 4096 + n * (w * h * 3 + 64)
        if out is None:
            out = np.empty(cap, dtype=np.uint8)
        ln, ne = C.c_uint64(), C.c_uint32()
        self._ck(self.lib.agmvb_encode_sequence(self.h, C.c_void_p(src), on_dev, n, w, h, create_n, fps, opt, quality,
                                                compression, C.c_void_p(out.ctypes.data), out.size, C.byref(ln), C.byref(ne)))
        return out[:ln.value], ne.value

    def encode_sequence_multi(self, others, frames, create_n, fps, opt, quality, compression=LZSS):
        """agmvb_encode_sequence_multi: this context plus `others` (one Context per further GPU), host frames (n,h,w) uint32."""
        frames = np.ascontiguousarray(frames, dtype=np.uint32)
        n, h, w = frames.shape
        out = np.empty(4096 + n * (w * h * 3 + 64), dtype=np.uint8)
        ctxs = (C.c_void_p * (1 + len(others)))(self.h, *[o.h for o in others])
        ln, ne = C.c_uint64(), C.c_uint32()
        self._ck(self.lib.agmvb_encode_sequence_multi(ctxs, len(ctxs), C.c_void_p(frames.ctypes.data), n, w, h, create_n, fps, opt, quality,
                                                      compression, C.c_void_p(out.ctypes.data), out.size, C.byref(ln), C.byref(ne)))
        return out[:ln.value], ne.value

    def encode_mode(self, mode, frames, create_n, fps, opt, quality, compression=LZSS):
        """mode 'video' = AGMV_EncodeVideo, 'full' = AGMV_EncodeFullAGMV, on host frames (n,h,w) uint32."""
        frames = np.ascontiguousarray(frames, dtype=np.uint32)
        n, h, w = frames.shape
        out = np.empty(4096 + n * (w * h * 3 + 64), dtype=np.uint8)
        ln, ne = C.c_uint64(), C.c_uint32()
        if mode == "video":
            rc = self.lib.agmvb_encode_video(self.h, C.c_void_p(frames.ctypes.data), 0, n, w, h, fps, opt, quality, compression,
                                             C.c_void_p(out.ctypes.data), out.size, C.byref(ln), C.byref(ne))
        else:
            rc = self.lib.agmvb_encode_full(self.h, C.c_void_p(frames.ctypes.data), 0, n, w, h, create_n, fps, opt, quality, compression,
                                            C.c_void_p(out.ctypes.data), out.size, C.byref(ln), C.byref(ne))
        self._ck(rc)
        return out[:ln.value], ne.value

    def enc_begin(self, w, h, opt, quality, compression=LZSS):
        self._ck(self.lib.agmvb_enc_begin(self.h, w, h, opt, quality, compression))

    def enc_histogram(self, ptr, n_frames, on_device):
        self._ck(self.lib.agmvb_enc_histogram(self.h, C.c_void_p(ptr), n_frames, 1 if on_device else 0))

    def enc_histogram_ptr(self):
        p, n = C.c_void_p(), C.c_uint32()
        self._ck(self.lib.agmvb_enc_histogram_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def enc_build_palette(self):
        self._ck(self.lib.agmvb_enc_build_palette(self.h))

    def enc_get_palette(self):
        p0, p1 = np.zeros(256, np.uint32), np.zeros(256, np.uint32)
        self._ck(self.lib.agmvb_enc_get_palette(self.h, _p(p0, _u32p), _p(p1, _u32p)))
        return p0, p1

    def enc_set_palette(self, p0, p1):
        p0 = np.ascontiguousarray(p0, np.uint32)
        p1 = np.ascontiguousarray(p1, np.uint32)
        self._ck(self.lib.agmvb_enc_set_palette(self.h, _p(p0, _u32p), _p(p1, _u32p)))

    def enc_frames(self, ptr, n_in_buffer, on_device, src_a, src_b, first_frame_count=0):
        sa = np.ascontiguousarray(src_a, np.int32)
        sb = np.ascontiguousarray(src_b, np.int32)
        nbytes = C.c_uint64()
        self._ck(self.lib.agmvb_enc_frames(self.h, C.c_void_p(ptr), n_in_buffer, 1 if on_device else 0, _p(sa, _i32p), _p(sb, _i32p),
                                           len(sa), first_frame_count, C.byref(nbytes)))
        return nbytes.value

    def enc_fetch(self, nbytes, n_enc, host_ptr=None):
        us, cs = np.zeros(n_enc, np.uint32), np.zeros(n_enc, np.uint32)
        img = None
        if host_ptr is None:
            img = np.empty(nbytes, np.uint8)
            host_ptr = img.ctypes.data
        self._ck(self.lib.agmvb_enc_fetch(self.h, C.c_void_p(host_ptr), nbytes, _p(us, _u32p), _p(cs, _u32p)))
        return img, us, cs

    def enc_sizes(self, n_enc):
        """usize / csize per frame of the last enc_frames call."""
        us, cs = np.zeros(n_enc, np.uint32), np.zeros(n_enc, np.uint32)
        self._ck(self.lib.agmvb_enc_fetch(self.h, None, 0, _p(us, _u32p), _p(cs, _u32p)))
        return us, cs

    def enc_image_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.agmvb_enc_image_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def enc_header(self, n_frames, fps):
        out = np.zeros(2048, np.uint8)
        ln = C.c_uint64()
        self._ck(self.lib.agmvb_enc_header(self.h, n_frames, fps, _p(out, _u8p), out.size, C.byref(ln)))
        return out[:ln.value]

    # ---- decoder ------------------------------------------------------------
    def dec_open(self, data):
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        sid, w, h, n = C.c_int(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._ck(self.lib.agmvb_dec_open(self.h, C.c_void_p(buf.ctypes.data), buf.size, C.byref(sid), C.byref(w), C.byref(h), C.byref(n)))
        return sid.value, w.value, h.value, n.value

    def dec_frames(self, sid, count, w, h, device_ptr=None, host_ptr=None):
        if device_ptr is not None:
            self._ck(self.lib.agmvb_dec_frames(self.h, sid, count, C.c_void_p(device_ptr), 1))
            return None
        out = None
        if host_ptr is None:
            out = np.empty((count, h, w, 3), dtype=np.uint8) if getattr(self, "host_fmt", 0) == 1 else np.empty((count, h, w), dtype=np.uint32)
            host_ptr = out.ctypes.data
        self._ck(self.lib.agmvb_dec_frames(self.h, sid, count, C.c_void_p(host_ptr), 0))
        return out

    def dec_batch(self, sids, count, out_ptrs=None, checksums=False):
        ids = (C.c_int * len(sids))(*sids)
        outs = None
        if out_ptrs is not None:
            outs = (C.c_void_p * len(sids))(*out_ptrs)
        ck = np.zeros((len(sids), count), np.uint64) if checksums else None
        self._ck(self.lib.agmvb_dec_batch(self.h, ids, len(sids), count, outs, _p(ck, _u64p) if checksums else None))
        return ck

    def dec_seek(self, sid, frame_index):
        """AGMV_SkipTo without I-frame rounding: the next frame decoded is frame_index; decoder state is not rewound."""
        self._ck(self.lib.agmvb_dec_seek(self.h, sid, frame_index))

    def dec_close(self, sid):
        self._ck(self.lib.agmvb_dec_close(self.h, sid))

    # ---- audio chunk codec ---------------------------------------
    def enc_set_audio(self, pcm, sample_rate, channels, total_duration):
        """The handle's audio track (what AGMV_WavToAudioTrack sets): uint16 or uint8 samples. None clears it."""
        if pcm is None:
            self._ck(self.lib.agmvb_enc_set_audio(self.h, None, 0, 16, 0, 0, 0))
            return
        pcm = np.ascontiguousarray(pcm)
        bits = {np.dtype(np.uint16): 16, np.dtype(np.uint8): 8}[pcm.dtype]
        self._ck(self.lib.agmvb_enc_set_audio(self.h, C.c_void_p(pcm.ctypes.data), pcm.size, bits, sample_rate, channels, total_duration))

    def enc_set_audio_chunk(self, chunk_size):
        self._ck(self.lib.agmvb_enc_set_audio_chunk(self.h, chunk_size))

    def enc_get_atsample(self, n):
        out = np.empty(n, np.uint8)
        self._ck(self.lib.agmvb_enc_get_atsample(self.h, _p(out, _u8p), n))
        return out

    def audio_compress(self, pcm):
        pcm = np.ascontiguousarray(pcm)
        bits = {np.dtype(np.uint16): 16, np.dtype(np.uint8): 8}[pcm.dtype]
        out = np.empty(pcm.size, np.uint8)
        self._ck(self.lib.agmvb_audio_compress(self.h, C.c_void_p(pcm.ctypes.data), pcm.size, bits, _p(out, _u8p)))
        return out

    def audio_expand(self, atsample, bits=16):
        atsample = np.ascontiguousarray(atsample, dtype=np.uint8)
        out = np.empty(atsample.size, np.uint16 if bits == 16 else np.uint8)
        self._ck(self.lib.agmvb_audio_expand(self.h, _p(atsample, _u8p), atsample.size, bits, C.c_void_p(out.ctypes.data)))
        return out

    def dec_audio(self, sid):
        """Every audio chunk of an open stream, decoded (AGMV_DecodeAudioChunk): uint16 or uint8 samples."""
        n, bits = C.c_uint64(), C.c_int()
        self._ck(self.lib.agmvb_dec_audio(self.h, sid, None, 0, C.byref(n), C.byref(bits)))
        out = np.empty(n.value, np.uint16 if bits.value == 16 else np.uint8)
        if n.value:
            self._ck(self.lib.agmvb_dec_audio(self.h, sid, C.c_void_p(out.ctypes.data), n.value, C.byref(n), C.byref(bits)))
        return out

    def decode_all(self, data):
        sid, w, h, n = self.dec_open(data)
        try:
            return self.dec_frames(sid, n, w, h)
        finally:
            self.dec_close(sid)

    # ---- measurement utilities ----------------------------------------------------
    def synth_frames(self, device_ptr, w, h, first_t, n, seed=1234):
        self._ck(self.lib.agmvb_synth_frames(self.h, C.c_void_p(device_ptr), w, h, first_t, n, seed))

    def profile(self, enable):
        self._ck(self.lib.agmvb_profile(self.h, 1 if enable else 0))

    def profile_read(self):
        n = self.lib.agmvb_profile_classes()
        cnt = np.zeros(n, np.uint64)
        ms = np.zeros(n, np.float64)
        self._ck(self.lib.agmvb_profile_read(self.h, _p(cnt, _u64p), _p(ms, C.POINTER(C.c_double))))
        return {self.lib.agmvb_profile_name(k).decode(): (int(cnt[k]), float(ms[k])) for k in range(n) if cnt[k]}

    # ---- unit-test hooks -------------------------------------------------------
    def test_peek(self, which, dtype, count):
        out = np.zeros(count, dtype=dtype)
        self._ck(self.lib.agmvb_test_peek(self.h, which, C.c_void_p(out.ctypes.data), out.nbytes))
        return out

    def test_lzss(self, buffers):
        """buffers: list of bytes/uint8 arrays -> list of (csize, bytes, outbits) as AGMV_LZSS would produce."""
        arrs = [np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b for b in buffers]
        fs = np.zeros(len(arrs) + 1, np.uint32)
        fs[1:] = np.cumsum([len(a) for a in arrs])
        data = np.concatenate(arrs + [np.zeros(1, np.uint8)]).astype(np.uint8)
        cap = int(fs[-1]) * 9 // 8 + 16 * len(arrs) + 64
        out = np.zeros(cap, np.uint8)
        off = np.zeros(len(arrs) + 1, np.uint64)
        cs, ob = np.zeros(len(arrs), np.uint32), np.zeros(len(arrs), np.uint32)
        self._ck(self.lib.agmvb_test_lzss(self.h, _p(data, _u8p), _p(fs, _u32p), len(arrs), _p(out, _u8p), cap, _p(off, _u64p),
                                          _p(cs, _u32p), _p(ob, _u32p)))
        return [(int(cs[k]), out[int(off[k]):int(off[k + 1])].tobytes(), int(ob[k])) for k in range(len(arrs))]

    def test_lz77(self, buffers, persist_fill=0):
        """buffers: consecutive frames' bitstreams of one handle -> list of (csize, token bytes) as AGMV_LZ77 would produce."""
        arrs = [np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b for b in buffers]
        fs = np.zeros(len(arrs) + 1, np.uint32)
        fs[1:] = np.cumsum([len(a) for a in arrs])
        data = np.concatenate(arrs + [np.zeros(1, np.uint8)]).astype(np.uint8)
        cap = int(fs[-1]) * 4 + 64
        out = np.zeros(cap, np.uint8)
        off = np.zeros(len(arrs) + 1, np.uint64)
        cs = np.zeros(len(arrs), np.uint32)
        self._ck(self.lib.agmvb_test_lz77(self.h, _p(data, _u8p), _p(fs, _u32p), len(arrs), persist_fill, _p(out, _u8p), cap, _p(off, _u64p),
                                          _p(cs, _u32p)))
        return [(int(cs[k]), out[int(off[k]):int(off[k + 1])].tobytes()) for k in range(len(arrs))]

    def test_quantize(self, colors, pal0, pal1, dual, full_table=False):
        colors = np.ascontiguousarray(colors, np.uint32)
        pal0 = np.ascontiguousarray(pal0, np.uint32)
        pal1 = np.ascontiguousarray(pal1, np.uint32)
        ent = np.zeros(colors.size, np.uint16)
        self._ck(self.lib.agmvb_test_quantize(self.h, _p(colors, _u32p), colors.size, _p(pal0, _u32p), _p(pal1, _u32p), (1 if dual else 0) | (2 if full_table else 0),
                                              _p(ent, _u16p)))
        return ent

    def test_assemble(self, entries, ientries, w, h, dual, pal0, pal1):
        entries = np.ascontiguousarray(entries, np.uint16)
        pal0 = np.ascontiguousarray(pal0, np.uint32)
        pal1 = np.ascontiguousarray(pal1, np.uint32)
        out = np.zeros(w * h * 33 // 16 + 64, np.uint8)
        us = C.c_uint32()
        ip = None
        if ientries is not None:
            ientries = np.ascontiguousarray(ientries, np.uint16)
            ip = _p(ientries, _u16p)
        self._ck(self.lib.agmvb_test_assemble(self.h, _p(entries, _u16p), ip, w, h, 1 if dual else 0, _p(pal0, _u32p), _p(pal1, _u32p),
                                              _p(out, _u8p), out.size, C.byref(us)))
        return out[:us.value].tobytes()
