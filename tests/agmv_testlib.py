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


def ref_encode(frames, create_n, fps, opt, quality, compression=LZSS, workdir=None, timing=None, mode="agmv"):
    """Run the reference encoder on BMP files of `frames` (frames 1..n). Returns .agmv bytes."""
    n, h, w = frames.shape
    with tempfile.TemporaryDirectory(dir=workdir) as td:
        # the reference formats paths into char[60] (src/agmv_encode.c:2372): keep them short
        write_bmps(frames, td, "f", 1)
        res = subprocess.run([os.path.join(REF_DIR, "ref_encode"), "o.agmv", ".", "f", "1", str(n), str(w), str(h),
                              str(fps), str(opt), str(quality), str(compression), str(create_n), mode],
                             cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
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
