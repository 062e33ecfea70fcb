"""Shared helpers for the test-suite, bench.py's CPU legs and smoke().

Everything under oracle/ is TEST INFRASTRUCTURE: this module is the only place
that loads it. The product path (libagmv_b200) never imports this file.
"""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

OPT = dict(I=1, II=2, III=3, ANIM=4, GBA_I=5, GBA_II=6, GBA_III=7, NDS=8)
QUALITY = dict(HIGH=1, MID=2, LOW=3)
LZSS, LZ77 = 1, 2

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


_oracle = None


def oracle():
    """ctypes handle on oracle/libagmv_oracle.so (our C restatement)."""
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libagmv_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.orc_synth_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, _u32p]
        lib.orc_write_bmp24.argtypes = [C.c_char_p, C.c_int, C.c_int, _u32p]
        lib.orc_max_clr.restype = C.c_uint32
        lib.orc_quantize_color.restype = C.c_uint32
        lib.orc_quantize_color.argtypes = [C.c_uint32, C.c_int]
        lib.orc_histogram_add.argtypes = [_u64p, _u32p, C.c_size_t, C.c_int]
        lib.orc_build_palette.argtypes = [_u64p, C.c_int, C.c_int, _u32p, _u32p]
        lib.orc_interp.argtypes = [_u32p, _u32p, _u32p, C.c_size_t]
        lib.orc_scale_nearest.argtypes = [_u32p, C.c_int, C.c_int, C.c_float, C.c_float, _u32p,
                                          C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.orc_nearest_entry.restype = C.c_uint16
        lib.orc_nearest_entry.argtypes = [_u32p, _u32p, C.c_int, C.c_uint32]
        lib.orc_quantize_frame.argtypes = [_u32p, C.c_size_t, _u32p, _u32p, C.c_int, _u16p]
        lib.orc_assemble.restype = C.c_size_t
        lib.orc_assemble.argtypes = [_u16p, _u16p, C.c_int, C.c_int, C.c_int, C.c_int, _u32p, _u32p, _u8p]
        lib.orc_lzss.restype = C.c_uint32
        lib.orc_lzss.argtypes = [_u8p, C.c_size_t, _u8p, C.POINTER(C.c_size_t), _u64p]
        lib.orc_encode_agmv.restype = C.c_long
        lib.orc_encode_agmv.argtypes = [_u32p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                        C.c_int, C.c_int, C.c_int, _u8p, C.c_size_t]
        lib.orc_decode_agmv.restype = C.c_int
        lib.orc_decode_agmv.argtypes = [_u8p, C.c_size_t, _u32p, C.c_size_t,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.orc_audio_compress16.argtypes = [_u16p, C.c_size_t, _u8p]
        lib.orc_audio_expand16.argtypes = [_u8p, C.c_size_t, _u16p]
        _oracle = lib
    return _oracle


def ptr(a, t):
    return a.ctypes.data_as(t)


def synth_frames(w, h, n, seed=1234, first=1):
    """n synthetic frames t=first..first+n-1 as an (n,h,w) uint32 0x00RRGGBB array."""
    out = np.empty((n, h, w), dtype=np.uint32)
    lib = oracle()
    for k in range(n):
        lib.orc_synth_frame(w, h, first + k, seed, ptr(out[k], _u32p))
    return out


def write_bmps(frames, directory, base="f", first=1):
    lib = oracle()
    n, h, w = frames.shape
    for k in range(n):
        fr = np.ascontiguousarray(frames[k])
        rc = lib.orc_write_bmp24(os.path.join(directory, f"{base}{first + k}.bmp").encode(), w, h, ptr(fr, _u32p))
        assert rc == 0


def oracle_encode(frames, create_n, fps, opt, quality, compression=LZSS):
    lib = oracle()
    n, h, w = frames.shape
    frames = np.ascontiguousarray(frames, dtype=np.uint32)
    cap = 2048 + n * (w * h * 3 + 64)
    out = np.zeros(cap, dtype=np.uint8)
    ln = lib.orc_encode_agmv(ptr(frames, _u32p), n, w, h, create_n, fps, opt, quality, compression,
                             ptr(out, _u8p), cap)
    assert ln > 0, f"oracle encode failed rc={ln}"
    return out[:ln].tobytes()


def oracle_encode_mode(mode, frames, create_n, fps, opt, quality, compression=LZSS):
    """mode 'video' = AGMV_EncodeVideo, 'full' = AGMV_EncodeFullAGMV."""
    lib = oracle()
    n, h, w = frames.shape
    frames = np.ascontiguousarray(frames, dtype=np.uint32)
    cap = 2048 + n * (w * h * 3 + 64)
    out = np.zeros(cap, dtype=np.uint8)
    if mode == "video":
        lib.orc_encode_video.restype = C.c_long
        lib.orc_encode_video.argtypes = [_u32p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, _u8p, C.c_size_t]
        ln = lib.orc_encode_video(ptr(frames, _u32p), n, w, h, fps, opt, quality, compression, ptr(out, _u8p), cap)
    else:
        lib.orc_encode_full.restype = C.c_long
        lib.orc_encode_full.argtypes = [_u32p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, _u8p, C.c_size_t]
        ln = lib.orc_encode_full(ptr(frames, _u32p), n, w, h, create_n, fps, opt, quality, compression, ptr(out, _u8p), cap)
    assert ln > 0, f"oracle encode failed rc={ln}"
    return out[:ln].tobytes()


def scene_cut_frames(w, h, n, every=5, seed=1234):
    """Synthetic frames with a scene cut (new tile colours) every `every` frames: consecutive frames inside a scene are
    ~2/3 grey-equal, across a cut ~0, so AGMV_EncodeVideo's similarity gate takes both branches."""
    out = np.empty((n, h, w), dtype=np.uint32)
    lib = oracle()
    for k in range(n):
        lib.orc_synth_frame(w, h, k + 1, seed + 7 * (k // every), ptr(out[k], _u32p))
    return out


def oracle_decode(data):
    lib = oracle()
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    w, h, n = C.c_int(), C.c_int(), C.c_int()
    rc = lib.orc_decode_agmv(ptr(buf, _u8p), len(buf), None, 0, C.byref(w), C.byref(h), C.byref(n))
    if rc != 0:
        return rc, None
    out = np.zeros((n.value, h.value, w.value), dtype=np.uint32)
    rc = lib.orc_decode_agmv(ptr(buf, _u8p), len(buf), ptr(out, _u32p), out.size, C.byref(w), C.byref(h), C.byref(n))
    return rc, out


def oracle_lzss(data):
    lib = oracle()
    src = np.frombuffer(data, dtype=np.uint8).copy() if not isinstance(data, np.ndarray) else data
    out = np.zeros(len(src) * 9 // 8 + 16, dtype=np.uint8)
    nb, bits = C.c_size_t(), C.c_uint64()
    csize = lib.orc_lzss(ptr(src, _u8p), len(src), ptr(out, _u8p), C.byref(nb), C.byref(bits))
    return csize, out[:nb.value].tobytes(), bits.value


# --------------------------------------------------------------------------
# the unmodified reference, one operation per fresh process
# --------------------------------------------------------------------------
def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "ref_encode")) and os.path.exists(os.path.join(REF_DIR, "ref_decode"))


def ref_encode(frames, create_n, fps, opt, quality, compression=LZSS, workdir=None, timing=None, mode="agmv", audio=None):
    """Run the reference encoder on BMP files of `frames` (frames 1..n). Returns .agmv bytes.
    audio = (pcm, sample_rate, channels): written as a WAV file and loaded with AGMV_WavToAudioTrack first."""
    n, h, w = frames.shape
    with tempfile.TemporaryDirectory(dir=workdir) as td:
        # the reference formats paths into char[60] (src/agmv_encode.c:2372): keep them short
        write_bmps(frames, td, "f", 1)
        env = dict(os.environ)
        if audio is not None:
            write_wav(os.path.join(td, "a.wav"), *audio)
            env["REF_WAV"] = "a.wav"
        res = subprocess.run([os.path.join(REF_DIR, "ref_encode"), "o.agmv", ".", "f", "1", str(n), str(w), str(h),
                              str(fps), str(opt), str(quality), str(compression), str(create_n), mode],
                             cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
        if timing is not None:
            for line in res.stderr.decode().splitlines():
                if line.startswith("ref_encode_seconds"):
                    timing["seconds"] = float(line.split()[1])
        with open(os.path.join(td, "o.agmv"), "rb") as f:
            return f.read()


def ref_decode_raw(data, timing=None):
    """Reference per-frame decode without BMP export. Returns (rc, frames or None)."""
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "i.agmv"), "wb") as f:
            f.write(data)
        res = subprocess.run([os.path.join(REF_DIR, "ref_decode"), "raw", "i.agmv", "o.raw"], cwd=td, check=True,
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if timing is not None:
            for line in res.stderr.decode().splitlines():
                if line.startswith("ref_decode_seconds"):
                    timing["seconds"] = float(line.split()[1])
        tok = res.stdout.decode().split()
        rc = int(tok[1])
        if len(tok) < 8:
            return rc, None
        n, w, h = int(tok[3]), int(tok[5]), int(tok[7])
        raw = np.fromfile(os.path.join(td, "o.raw"), dtype=np.uint32)
        return rc, raw.reshape(n, h, w)


def sha256(b):
    return hashlib.sha256(b).hexdigest()


def oracle_lz77(data, stale=0xA7):
    """orc_lz77 on `data`; the byte one past the end (read when the last match ends at the buffer end) is `stale`."""
    lib = oracle()
    lib.orc_lz77.restype = C.c_uint32
    lib.orc_lz77.argtypes = [_u8p, C.c_size_t, _u8p, C.POINTER(C.c_size_t)]
    src = np.concatenate([np.asarray(data, dtype=np.uint8), np.full(8, stale, dtype=np.uint8)])
    out = np.zeros(len(data) * 4 + 16, dtype=np.uint8)
    nb = C.c_size_t()
    csize = lib.orc_lz77(ptr(src, _u8p), len(data), ptr(out, _u8p), C.byref(nb))
    return csize, out[:nb.value].tobytes()


def ref_decode_seek(data, plan):
    """Reference decode visiting frames in the order of `plan` (oracle/ref_decode.c seek mode). Returns (rc, frames)."""
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "i.agmv"), "wb") as f:
            f.write(data)
        res = subprocess.run([os.path.join(REF_DIR, "ref_decode"), "seek", "i.agmv", "o.raw", ",".join(str(k) for k in plan)], cwd=td,
                             check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        tok = res.stdout.decode().split()
        rc, n, w, h = int(tok[1]), int(tok[3]), int(tok[5]), int(tok[7])
        raw = np.fromfile(os.path.join(td, "o.raw"), dtype=np.uint32)
        return rc, raw.reshape(n, h, w)


# ---- audio chunk codec ---------------------------------------------------------------
def oracle_audio_compress16(pcm):
    pcm = np.ascontiguousarray(pcm, dtype=np.uint16)
    out = np.empty(pcm.size, np.uint8)
    oracle().orc_audio_compress16(ptr(pcm, _u16p), pcm.size, ptr(out, _u8p))
    return out


def oracle_audio_expand16(at):
    at = np.ascontiguousarray(at, dtype=np.uint8)
    out = np.empty(at.size, np.uint16)
    oracle().orc_audio_expand16(ptr(at, _u8p), at.size, ptr(out, _u16p))
    return out


def synth_pcm(n, bits=16, seed=7):
    """Deterministic synthetic track: a swept tone around mid-scale, a quiet noise floor, bursts reaching both rails."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64)
    x = 0.6 * np.sin(t * (0.01 + t / n * 0.2)) + 0.05 * rng.standard_normal(n)
    x[n // 3:n // 3 + n // 20] *= 3.0
    full = (1 << bits) - 1
    v = np.clip(np.rint((x * 0.5 + 0.5) * full), 0, full)
    v[:min(n, 8)] = [0, 1, full, full - 1, 255, 256, 65280 & full, 65281 & full][:min(n, 8)]
    return v.astype(np.uint16 if bits == 16 else np.uint8)


def write_wav(path, pcm, sample_rate, channels):
    """The 44-byte header AGMV_WavToAudioTrack reads (src/agmv_utils.c:1043-1056). The reference takes the RIFF chunk size
    at offset 4 as the number of sample BYTES (audio_size = chunk_size / 2 for 16-bit), so that field is written as
    exactly the data size - with the usual 36 + data the reference's track would end in 18 uninitialised samples."""
    pcm = np.ascontiguousarray(pcm)
    bits = pcm.dtype.itemsize * 8
    data = pcm.tobytes()
    import struct
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, channels, sample_rate,
                sample_rate * channels * bits // 8, channels * bits // 8, bits) + b"data" + struct.pack("<I", len(data)))
        f.write(data)


def audio_duration(pcm, sample_rate, channels):
    """header.total_audio_duration as AGMV_WavToAudioTrack computes it (src/agmv_utils.c:1070)."""
    return pcm.nbytes // (sample_rate * channels * pcm.dtype.itemsize)


def ref_audio(kind, arr):
    """kind 'compress': uint16 samples -> bytes through the reference's AGMV_CompressAudio; 'expand': the reverse
    through AGMV_DecodeAudioChunk."""
    with tempfile.TemporaryDirectory() as td:
        np.ascontiguousarray(arr).tofile(os.path.join(td, "i.bin"))
        subprocess.run([os.path.join(REF_DIR, "ref_audio"), kind, "i.bin", "o.bin"], cwd=td, check=True, stdout=subprocess.DEVNULL)
        return np.fromfile(os.path.join(td, "o.bin"), dtype=np.uint8 if kind == "compress" else np.uint16)


def ref_audio_track(data):
    """The reference's decode loop for a stream with audio (frames + audio chunks). Returns (rc, samples)."""
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, "i.agmv"), "wb") as f:
            f.write(data)
        res = subprocess.run([os.path.join(REF_DIR, "ref_audio"), "track", "i.agmv", "o.bin"], cwd=td, check=True, stdout=subprocess.PIPE)
        tok = res.stdout.decode().split()
        rc = int(tok[1])
        if len(tok) < 6:
            return rc, None
        return rc, np.fromfile(os.path.join(td, "o.bin"), dtype=np.uint16 if int(tok[5]) == 16 else np.uint8)


def audio_chunk_size(audio_size, n_src, opt, mode="agmv"):
    """audio_chunk->size = audio_size / (f32)frames (src/agmv_encode.c:2661-2663 adjusted count, :4024-4025 end - start)."""
    adj = n_src - 1
    if mode == "agmv":
        if opt in (OPT["I"], OPT["ANIM"], OPT["GBA_I"], OPT["GBA_II"]):
            adj //= 2
        elif opt == OPT["GBA_III"]:
            adj = int(np.float32(adj) * np.float32(0.75))
        else:
            adj = int(adj * 0.75)
    return int(np.float32(audio_size) / np.float32(adj))


def mux_audio(stream, atsample, chunk, pcm, sample_rate, channels, mode="agmv"):
    """Expected stream with an audio track, built from the same stream without one: header audio fields
    (src/agmv_encode.c:41-49) and 'AGAC' chunk | chunk bytes after every frame chunk (AGMV_EncodeAudioChunk, :707-717),
    frame g taking atsample[g*chunk : (g+1)*chunk]."""
    import struct
    s = bytes(stream)
    dual = s[17] in (1, 3)
    o = 38 + 768 * (2 if dual else 1)
    out = bytearray(s[:o])
    out[22:38] = struct.pack("<IIIHH", audio_duration(pcm, sample_rate, channels), sample_rate, pcm.size, channels, pcm.dtype.itemsize * 8)
    g = 0
    while o < len(s):
        assert s[o:o + 4] == b"AGFC"
        ln = 24 + int.from_bytes(s[o + 12:o + 16], "little")
        out += s[o:o + ln]
        o += ln
        if mode == "agmv":
            assert s[o:o + 8] == b"AGAC\0\0\0\0"
            o += 8
        piece = bytes(atsample[g * chunk:(g + 1) * chunk])
        out += b"AGAC" + struct.pack("<I", chunk) + piece + bytes(chunk - len(piece))  # past the track: defined as 0 (the reference over-reads)
        g += 1
    return bytes(out)
