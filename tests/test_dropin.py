"""The reference-facing C API (include/agmv_dropin.h, libagmv_dropin.so).

CPU part: handle layout is binary compatible with the reference's headers.
GPU part (-m gpu): CreateAGMV / AGMV_EncodeAGMV / AGMV_DecodeAGMV / AGMV_EncodeFrame / AGMV_DecodeFrameChunk
called the way the reference's examples call them, results compared with the reference's golden output.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import libagmv_b200
from agmv_testlib import GOLDEN_DIR, LZSS, OPT, QUALITY, REF_DIR, have_ref, sha256, synth_frames, write_bmps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UL = C.c_ulong


class Hdr(C.Structure):
    _fields_ = [("fourcc", C.c_char * 4), ("num_of_frames", UL), ("width", UL), ("height", UL), ("fmt", C.c_uint8), ("version", C.c_uint8),
                ("frames_per_second", UL), ("total_audio_duration", UL), ("sample_rate", UL), ("audio_size", UL),
                ("num_of_channels", C.c_uint16), ("bits_per_sample", C.c_uint16), ("palette0", UL * 256), ("palette1", UL * 256)]


class Frame(C.Structure):
    _fields_ = [("width", UL), ("height", UL), ("img_data", C.POINTER(UL))]


class Bitstream(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", UL), ("pos", UL)]


class Entry(C.Structure):
    _fields_ = [("pal_num", C.c_uint8), ("index", C.c_uint8), ("occurence", UL)]


class AGMV(C.Structure):
    _fields_ = [("header", Hdr), ("frame_chunk", C.c_void_p), ("audio_chunk", C.c_void_p), ("bitstream", C.POINTER(Bitstream)),
                ("frame", C.POINTER(Frame)), ("iframe", C.POINTER(Frame)), ("audio_track", C.c_void_p), ("iframe_entries", C.POINTER(Entry)),
                ("opt", C.c_int), ("compression", C.c_int), ("frame_count", UL), ("leniency", C.c_float), ("offset_table", UL * 40000),
                ("enable_audio", C.c_int), ("volume", C.c_float)]


def test_handle_layout_matches_reference_abi():
    """sizeof / offsetof through a C compiler: our header vs the numbers measured on the reference."""
    src = r'''
    #include <stddef.h>
    #include <stdio.h>
    #include "INCLUDE_ME"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(AGMV), sizeof(AGMV_ENTRY), sizeof(u32), offsetof(AGMV, frame_count),
                           offsetof(AGMV, iframe_entries), offsetof(AGMV_MAIN_HEADER, palette0)); return 0; }
    '''
    def run(header, inc):
        with tempfile.TemporaryDirectory() as td:
            open(os.path.join(td, "t.c"), "w").write(src.replace("INCLUDE_ME", header))
            subprocess.run(["gcc", "-o", os.path.join(td, "t"), os.path.join(td, "t.c")] + [f"-I{i}" for i in inc], check=True)
            return subprocess.run([os.path.join(td, "t")], capture_output=True, text=True, check=True).stdout.split()
    ours = run("agmv_dropin.h", [os.path.join(ROOT, "include")])
    assert ours[:3] == ["324264", "16", "8"]
    assert C.sizeof(AGMV) == 324264 and C.sizeof(Entry) == 16
    if os.path.isdir("/root/reference/include"):
        theirs = run("agmv_defines.h", ["/root/reference/include"])
        assert ours == theirs


def _dropin():
    lib = C.CDLL(libagmv_b200.DROPIN_PATH)
    lib.CreateAGMV.restype = C.POINTER(AGMV)
    lib.CreateAGMV.argtypes = [UL, UL, UL, UL]
    lib.DestroyAGMV.argtypes = [C.POINTER(AGMV)]
    lib.AGMV_EncodeAGMV.restype = None
    lib.AGMV_EncodeAGMV.argtypes = [C.POINTER(AGMV), C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint8, UL, UL, UL, UL, UL, C.c_int, C.c_int, C.c_int]
    lib.AGMV_DecodeAGMV.restype = C.c_int
    lib.AGMV_DecodeAGMV.argtypes = [C.c_char_p, C.c_uint8, C.c_int]
    lib.AGMV_DecodeHeader.restype = C.c_int
    lib.AGMV_DecodeHeader.argtypes = [C.c_void_p, C.POINTER(AGMV)]
    lib.AGMV_DecodeFrameChunk.restype = C.c_int
    lib.AGMV_DecodeFrameChunk.argtypes = [C.c_void_p, C.POINTER(AGMV)]
    lib.AGMV_EncodeFrame.restype = None
    lib.AGMV_EncodeFrame.argtypes = [C.c_void_p, C.POINTER(AGMV), C.POINTER(UL)]
    lib.AGMV_EncodeVideo.restype = None
    lib.AGMV_EncodeVideo.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint8, UL, UL, UL, UL, UL, C.c_int, C.c_int, C.c_int]
    lib.AGMV_EncodeFullAGMV.restype = None
    lib.AGMV_EncodeFullAGMV.argtypes = lib.AGMV_EncodeAGMV.argtypes
    lib.AGMV_B200_LastError.restype = C.c_char_p
    return lib


def _libc():
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    libc.ftell.restype = C.c_long
    libc.ftell.argtypes = [C.c_void_p]
    libc.fseek.argtypes = [C.c_void_p, C.c_long, C.c_int]
    return libc


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["syn96x80_III_LOW", "gba240_GBA_I_LOW", "syn64_II_LOW", "lz77_64_II_LOW", "lz77_96x80_I_MID"])
def test_encode_agmv_dropin_bytes(golden, name):
    """examples/simple_video/simple_video.c, line for line: CreateAGMV + AGMV_EncodeAGMV on a BMP directory."""
    lz77 = name.startswith("lz77")
    g = dict(golden["encode_lz77" if lz77 else "encode"][name])
    g.setdefault("seed", 1234)
    comp = 2 if lz77 else LZSS
    lib = _dropin()
    frames = synth_frames(g["w"], g["h"], g["n"], seed=g["seed"])
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        write_bmps(frames, td, "f", 1)
        os.chdir(td)
        try:
            h = lib.CreateAGMV(g["create_n"], g["w"], g["h"], g["fps"])
            lib.AGMV_EncodeAGMV(h, b"o.agmv", b".", b"f", 1, 1, g["n"], g["w"], g["h"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], comp)
            data = open("o.agmv", "rb").read()
            if g["opt"].startswith("GBA"):
                assert os.path.exists("GBA_GEN_AGMV.h")
                txt = open("GBA_GEN_AGMV.h").read()
                assert f"GBA_AGMV_FILE[{len(data)}]" in txt
        finally:
            os.chdir(cwd)
    assert (len(data), sha256(data)) == (g["size"], g["sha256"]), lib.AGMV_B200_LastError()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["syn64_III_LOW", "lz77_64_II_LOW"])
def test_decode_agmv_dropin_exports(golden, name):
    """AGMV_DecodeAGMV writes ./quick_export_<k>.bmp; pixel content must equal the reference's frames and,
    where the reference binary is present, the files must be byte-identical to its own export."""
    g = golden["encode_lz77" if name.startswith("lz77") else "encode"][name]
    path = os.path.join(GOLDEN_DIR, g["file"])
    lib = _dropin()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td, tempfile.TemporaryDirectory() as td2:
        os.chdir(td)
        try:
            assert lib.AGMV_DecodeAGMV(path.encode(), 1, 1) == 0
            assert lib.AGMV_DecodeAGMV(b"does_not_exist.agmv", 1, 1) == 2
            bad = bytearray(open(path, "rb").read())
            bad[36:38] = (36904).to_bytes(2, "little")
            open("bad.agmv", "wb").write(bad)
            assert lib.AGMV_DecodeAGMV(b"bad.agmv", 1, 1) == 1
            files = sorted(os.listdir(td))
            bmps = [f for f in files if f.startswith("quick_export_")]
            assert len(bmps) == g["decoded_shape"][0]
            mine = {f: open(f, "rb").read() for f in bmps}
        finally:
            os.chdir(cwd)
        # the export counter is process-global in the reference (extern/agidl/src/agidl_img_export.c:18): take the files in order
        order = sorted(bmps, key=lambda f: int(f[len("quick_export_"):-4]))
        for k in range(len(bmps)):
            d = mine[order[k]]
            px = np.frombuffer(d[54:], dtype=np.uint8).reshape(g["h"], g["w"], 3).astype(np.uint32)
            frame = px[..., 2] << 16 | px[..., 1] << 8 | px[..., 0]
            assert sha256(frame.astype(np.uint32).tobytes()) == g["decoded_frame_sha256"][k]
        if have_ref():
            subprocess.run([os.path.join(REF_DIR, "ref_decode"), "export", path], cwd=td2, check=True, capture_output=True)
            for k, f in enumerate(order):
                assert open(os.path.join(td2, f"quick_export_{k + 1}.bmp"), "rb").read() == mine[f], f


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gba240_GBA_I_LOW", "lz77_gba240_GBA_I_LOW"])
def test_decode_frame_chunk_dropin_streaming(golden, name):
    """AGMV_PlayAGMV's loop (src/agmv_playback.c:102-115): find the next 'AGFC', AGMV_DecodeFrameChunk, repeat."""
    g = golden["encode_lz77" if name.startswith("lz77") else "encode"][name]
    path = os.path.join(GOLDEN_DIR, g["file"])
    raw = open(path, "rb").read()
    lib, libc = _dropin(), _libc()
    f = libc.fopen(path.encode(), b"rb")
    h = lib.CreateAGMV(0, 120, 80, 16)  # sizes are overwritten by the header
    assert lib.AGMV_DecodeHeader(f, h) == 0
    a = h.contents
    assert (a.header.width, a.header.height, a.header.num_of_frames) == (120, 80, g["decoded_shape"][0])
    a.frame.contents.width = a.iframe.contents.width = a.header.width
    a.frame.contents.height = a.iframe.contents.height = a.header.height
    a.frame_count = 0
    P = 120 * 80
    for k in range(a.header.num_of_frames):
        pos = libc.ftell(f)
        nxt = raw.find(b"AGFC", pos)  # AGMV_FindNextFrameChunk
        libc.fseek(f, nxt, 0)
        assert lib.AGMV_DecodeFrameChunk(f, h) == 0, lib.AGMV_B200_LastError()
        px = np.ctypeslib.as_array(a.frame.contents.img_data, shape=(P,)).astype(np.uint32)
        assert sha256(px.tobytes()) == g["decoded_frame_sha256"][k], k
        assert a.frame_count == k + 1
    libc.fclose(f)
    lib.DestroyAGMV(h)


@pytest.mark.gpu
@pytest.mark.parametrize("comp", [LZSS, 2])
def test_encode_frame_dropin_per_frame(golden, comp):
    """AGMV_EncodeFrame frame by frame (palettes taken from the reference's stream) reproduces the reference's chunks,
    with either entropy coder (LZ77 carries its bitstream buffer from frame to frame)."""
    g = dict(golden["encode"]["syn64_III_LOW"] if comp == LZSS else golden["encode_lz77"]["lz77_64_III_LOW"])
    g.setdefault("seed", 1234)
    ref = open(os.path.join(GOLDEN_DIR, g["file"]), "rb").read()
    lib, libc = _dropin(), _libc()
    frames = synth_frames(g["w"], g["h"], g["n"], seed=g["seed"]).reshape(g["n"], -1)
    h = lib.CreateAGMV(g["create_n"], g["w"], g["h"], g["fps"])
    a = h.contents
    a.opt, a.compression = OPT["III"], comp
    pal = np.frombuffer(ref[38:38 + 1536], dtype=np.uint8).reshape(512, 3).astype(np.uint64)
    for i in range(256):
        a.header.palette0[i] = int(pal[i, 0] << 16 | pal[i, 1] << 8 | pal[i, 2])
        a.header.palette1[i] = int(pal[256 + i, 0] << 16 | pal[256 + i, 1] << 8 | pal[256 + i, 2])
    # LIGHT schedule of the first two groups: f1, interp(f2,f3), f4, f5, interp(f6,f7), f8
    def interp(x, y):
        out = np.zeros_like(x)
        for sh in (16, 8, 0):
            c1, c2 = ((x >> sh) & 255).astype(np.int64), ((y >> sh) & 255).astype(np.int64)
            out |= ((c1 + ((c2 - c1) >> 1)).astype(np.uint32) & 255) << sh
        return out
    seq = [frames[0], interp(frames[1], frames[2]), frames[3], frames[4], interp(frames[5], frames[6]), frames[7]]
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "o.bin").encode()
        f = libc.fopen(p, b"wb")
        for img in seq:
            wide = img.astype(np.uint64)
            lib.AGMV_EncodeFrame(f, h, wide.ctypes.data_as(C.POINTER(UL)))
        libc.fclose(f)
        got = open(p, "rb").read()
    lib.DestroyAGMV(h)
    # the reference interleaves an 8-byte empty AGAC chunk after every frame chunk; strip those from its stream
    body = ref[38 + 1536:]
    exp, o = b"", 0
    for _ in range(6):
        cs = int.from_bytes(body[o + 12:o + 16], "little")
        exp += body[o:o + 16 + cs + 8]
        o += 16 + cs + 8 + 8
    assert got == exp


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["video64_III_LOW", "video240_GBA_I_LOW", "full64_III_LOW"])
def test_other_encoders_dropin_bytes(golden, name):
    """AGMV_EncodeVideo / AGMV_EncodeFullAGMV called like tools/agmvcli/agmvcli.c:168 does, on a BMP directory."""
    from agmv_testlib import scene_cut_frames
    g = golden["encode_modes"][name]
    lib = _dropin()
    frames = scene_cut_frames(g["w"], g["h"], g["n"])
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        write_bmps(frames, td, "f", 1)
        os.chdir(td)
        try:
            if g["mode"] == "video":
                lib.AGMV_EncodeVideo(b"o.agmv", b".", b"f", 1, 1, g["n"], g["w"], g["h"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZSS)
            else:
                h = lib.CreateAGMV(g["create_n"], g["w"], g["h"], g["fps"])
                lib.AGMV_EncodeFullAGMV(h, b"o.agmv", b".", b"f", 1, 1, g["n"], g["w"], g["h"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZSS)
            data = open("o.agmv", "rb").read()
        finally:
            os.chdir(cwd)
    assert (len(data), sha256(data)) == (g["size"], g["sha256"]), lib.AGMV_B200_LastError()


@pytest.mark.gpu
@pytest.mark.parametrize("fn", ["gba240_GBA_I_LOW.agmv", "lz77_320x240_III_LOW.agmv"])
def test_seek_dropin_like_skipto(golden, fn):
    """What AGMV_SkipTo does with the reference's own agmv_utils.o in the link (src/agmv_playback.c:94-100): fseek to the
    chunk, set frame_count, then AGMV_DecodeFrameChunk - served by the drop-in, frames equal to the reference's."""
    g = golden["seek"][fn]
    path = os.path.join(GOLDEN_DIR, fn)
    raw = open(path, "rb").read()
    offs, p = [], raw.find(b"AGFC")
    while p >= 0:   # AGMV_ParseAGMV's offset_table: one entry per 'AGFC'
        offs.append(p)
        cs = int.from_bytes(raw[p + 12:p + 16], "little")
        p = raw.find(b"AGFC", p + 16 + cs)
    lib, libc = _dropin(), _libc()
    f = libc.fopen(path.encode(), b"rb")
    h = lib.CreateAGMV(0, int.from_bytes(raw[8:12], "little"), int.from_bytes(raw[12:16], "little"), 16)  # buffers sized for the stream
    assert lib.AGMV_DecodeHeader(f, h) == 0
    a = h.contents
    a.frame.contents.width = a.iframe.contents.width = a.header.width
    a.frame.contents.height = a.iframe.contents.height = a.header.height
    a.frame_count = 0
    P = int(a.header.width * a.header.height)
    prev = -2
    for k, want in zip(g["plan"], g["frame_sha256"]):
        if k != prev + 1:
            libc.fseek(f, offs[k], 0)
            a.frame_count = k
        else:
            libc.fseek(f, raw.find(b"AGFC", libc.ftell(f)), 0)   # AGMV_FindNextFrameChunk
        assert lib.AGMV_DecodeFrameChunk(f, h) == 0, lib.AGMV_B200_LastError()
        px = np.ctypeslib.as_array(a.frame.contents.img_data, shape=(P,)).astype(np.uint32)
        assert sha256(px.tobytes()) == want, k
        prev = k
    libc.fclose(f)
    lib.DestroyAGMV(h)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["start3_64_III_LOW", "start2_64_I_LOW"])
def test_encode_agmv_dropin_frame_range(golden, name):
    """AGMV_EncodeAGMV with start_frame > 1: the BMP directory holds f1..f16, frames start..end are encoded."""
    g = golden["encode_ranges"][name]
    lib = _dropin()
    frames = synth_frames(g["w"], g["h"], g["n"], seed=g["seed"])
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        write_bmps(frames, td, "f", 1)
        os.chdir(td)
        try:
            h = lib.CreateAGMV(g["create_n"], g["w"], g["h"], g["fps"])
            lib.AGMV_EncodeAGMV(h, b"o.agmv", b".", b"f", 1, g["start"], g["end"], g["w"], g["h"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZSS)
            data = open("o.agmv", "rb").read()
        finally:
            os.chdir(cwd)
    assert (len(data), sha256(data)) == (g["size"], g["sha256"]), lib.AGMV_B200_LastError()


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="needs the reference sources (build container only)")
def test_integration_link_recipe():
    """INTEGRATION.md section 1, performed: the reference's own objects WITHOUT agmv_encode.o / agmv_decode.o - agmv_utils.c,
    agmv_playback.c, all of AGIDL and tools/agmvcli/agmvcli.c, compiled from /root/reference - link against
    -lagmv_dropin -lagmv_b200 with nothing left undefined (to_80bitfloat, AGMV_DecodeVideo, AGMV_DecodeAudio included)."""
    libdir = os.path.join(ROOT, "libagmv_b200")
    srcs = [os.path.join(REFERENCE, "tools/agmvcli/agmvcli.c"), os.path.join(REFERENCE, "src/agmv_utils.c"),
            os.path.join(REFERENCE, "src/agmv_playback.c")]
    agidl = os.path.join(REFERENCE, "extern/agidl/src")
    srcs += [os.path.join(agidl, f) for f in sorted(os.listdir(agidl)) if f.endswith(".c") and f != "main.c"]
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "agmvcli")
        r = subprocess.run(["gcc", "-O1", "-w", f"-I{REFERENCE}/extern/agidl/include", f"-I{REFERENCE}/include", "-o", exe] + srcs +
                           [f"-L{libdir}", "-lagmv_dropin", "-lagmv_b200", "-lm", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        # every hot-path symbol the CLI and agmv_utils.o need is resolved by the drop-in, none by a stray reference object
        und = subprocess.run(["nm", "-u", exe], capture_output=True, text=True, check=True).stdout
        for sym in ("AGMV_EncodeAGMV", "AGMV_EncodeVideo", "AGMV_DecodeAGMV", "AGMV_DecodeAudioChunk", "to_80bitfloat"):
            assert sym in und, f"{sym} is not taken from the shared drop-in"
        # usage text without a script: the binary loads both shared libraries and runs
        out = subprocess.run([exe], capture_output=True, text=True)
        assert "AGMVCLI" in out.stdout


def _agmvcli():
    exe = os.path.join(REF_DIR, "agmvcli_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/agmvcli_dropin not built (oracle/Makefile builds it where the reference sources are)")
    return exe


@pytest.mark.gpu
def test_agmvcli_on_dropin_encode_script(golden):
    """The reference's CLI (tools/agmvcli/agmvcli.c, unmodified, linked per INTEGRATION.md section 1) run on an .agvs script:
    '$video ENC ...' without an audio track calls AGMV_EncodeVideo (:168) - bytes equal the reference's golden stream."""
    from agmv_testlib import scene_cut_frames
    g = golden["encode_modes"]["video64_III_LOW"]
    exe = _agmvcli()
    frames = scene_cut_frames(g["w"], g["h"], g["n"])
    with tempfile.TemporaryDirectory() as td:
        write_bmps(frames, td, "f", 1)
        open(os.path.join(td, "enc.agvs"), "w").write(f"$video ENC o.agmv . f BMP 1 {g['n']} {g['w']} {g['h']} {g['fps']} OPT_III LOW_Q LZSS\n")
        r = subprocess.run([exe, "enc.agvs"], cwd=td, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        data = open(os.path.join(td, "o.agmv"), "rb").read()
    assert (len(data), sha256(data)) == (g["size"], g["sha256"])


@pytest.mark.gpu
def test_agmvcli_on_dropin_decode_script(golden):
    """'$video DEC cur <file> BMP WAV' -> AGMV_DecodeAGMV (:185-190): the exported frames equal the reference's."""
    g = golden["encode"]["syn64_III_LOW"]
    exe = _agmvcli()
    with tempfile.TemporaryDirectory() as td:
        import shutil
        shutil.copy(os.path.join(GOLDEN_DIR, g["file"]), os.path.join(td, "s.agmv"))
        open(os.path.join(td, "dec.agvs"), "w").write("$video DEC cur s.agmv BMP WAV\n")
        r = subprocess.run([exe, "dec.agvs"], cwd=td, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        bmps = sorted((f for f in os.listdir(td) if f.startswith("quick_export_")), key=lambda f: int(f[len("quick_export_"):-4]))
        assert len(bmps) == g["decoded_shape"][0]
        for k, f in enumerate(bmps):
            d = open(os.path.join(td, f), "rb").read()
            px = np.frombuffer(d[54:], dtype=np.uint8).reshape(g["h"], g["w"], 3).astype(np.uint32)
            frame = px[..., 2] << 16 | px[..., 1] << 8 | px[..., 0]
            assert sha256(frame.astype(np.uint32).tobytes()) == g["decoded_frame_sha256"][k]


@pytest.mark.gpu
def test_encode_agmv_dropin_on_the_reference_foxlogo_bmps(golden):
    """examples/simple_video/simple_video.c, served by the drop-in: CreateAGMV + AGMV_EncodeAGMV on a directory of the
    reference's own 24-bit BMP files (the first 20 frames of examples/simple_video/foxlogo, tests/golden/foxlogo/)."""
    g = golden["fixtures"]["foxlogo_I_LOW"]
    lib = _dropin()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.symlink(os.path.join(GOLDEN_DIR, g["dir"]), os.path.join(td, "fx"))   # the reference formats paths into char[60]
        os.chdir(td)
        try:
            h = lib.CreateAGMV(g["create_n"], g["w"], g["h"], g["fps"])
            lib.AGMV_EncodeAGMV(h, b"o.agmv", b"fx", g["base"].encode(), 1, 1, g["n"], g["w"], g["h"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZSS)
            data = open("o.agmv", "rb").read()
        finally:
            os.chdir(cwd)
    assert (len(data), sha256(data)) == (g["size"], g["sha256"]), lib.AGMV_B200_LastError()


class FrameChunk(C.Structure):
    _fields_ = [("fourcc", C.c_char * 4), ("frame_num", UL), ("uncompressed_size", UL), ("compressed_size", UL)]


def _play_visit(lib, libc, path, raw, offs, plan, ahead):
    """Visit the frames of a stream in `plan` order the way a player does (AGMV_SkipTo for a jump, AGMV_FindNextFrameChunk for the
    next frame, then AGMV_DecodeFrameChunk); returns per visit (frame sha256, bitstream->pos, file cursor) and the queue's counters."""
    os.environ["AGMV_B200_AHEAD"] = str(ahead)
    st0 = [C.c_ulong(), C.c_ulong(), C.c_ulong()]
    lib.AGMV_B200_PlayQueueStats(*[C.byref(x) for x in st0])
    f = libc.fopen(path.encode(), b"rb")
    h = lib.CreateAGMV(0, int.from_bytes(raw[8:12], "little"), int.from_bytes(raw[12:16], "little"), 16)
    assert lib.AGMV_DecodeHeader(f, h) == 0
    a = h.contents
    a.frame.contents.width = a.iframe.contents.width = a.header.width
    a.frame.contents.height = a.iframe.contents.height = a.header.height
    a.frame_count = 0
    P = int(a.header.width * a.header.height)
    out, prev = [], -2
    for k in plan:
        if k != prev + 1:
            libc.fseek(f, offs[k], 0)
            a.frame_count = k
        else:
            libc.fseek(f, raw.find(b"AGFC", libc.ftell(f)), 0)
        assert lib.AGMV_DecodeFrameChunk(f, h) == 0, lib.AGMV_B200_LastError()
        px = np.ctypeslib.as_array(a.frame.contents.img_data, shape=(P,)).astype(np.uint32)
        ipx = np.ctypeslib.as_array(a.iframe.contents.img_data, shape=(P,)).astype(np.uint32)
        fc = C.cast(a.frame_chunk, C.POINTER(FrameChunk)).contents   # include/agmv_defines.h: AGMV_FRAME_CHUNK
        out.append((sha256(px.tobytes()), sha256(ipx.tobytes()), int(a.bitstream.contents.pos), libc.ftell(f), int(a.frame_count),
                    bytes(fc.fourcc), int(fc.frame_num), int(fc.uncompressed_size), int(fc.compressed_size)))
        prev = k
    libc.fclose(f)
    lib.DestroyAGMV(h)
    st1 = [C.c_ulong(), C.c_ulong(), C.c_ulong()]
    lib.AGMV_B200_PlayQueueStats(*[C.byref(x) for x in st1])
    os.environ.pop("AGMV_B200_AHEAD", None)
    return out, [b.value - a_.value for a_, b in zip(st0, st1)]


@pytest.mark.gpu
@pytest.mark.parametrize("comp", [LZSS, 2])
def test_play_queue_is_invisible(tmp_path, comp):
    """the frame-ahead queue behind AGMV_DecodeFrameChunk (AGMV_PlayAGMV's loop, src/agmv_playback.c:102-115).
    A 60-frame stream is played straight through and then with jumps in the middle of batches (AGMV_SkipTo,
    src/agmv_playback.c:94-100); with the queue on (depth 8 and 5) every visit must leave exactly what the one-frame-per-call
    path leaves in the handle and the FILE - pixels, I-frame copy, bitstream->pos, cursor, frame_count, chunk fields - and
    the queue must actually have served frames and withdrawn batches."""
    lib, libc = _dropin(), _libc()
    lib.AGMV_B200_PlayQueueStats.argtypes = [C.POINTER(C.c_ulong)] * 3
    frames = synth_frames(96, 80, 80, seed=77)
    ctx = libagmv_b200.Context(0)
    data, n_enc = ctx.encode_sequence(frames, 79, 24, OPT["III"], QUALITY["LOW"], comp)
    ctx.close()
    raw = data.tobytes()
    path = str(tmp_path / "play.agmv")
    open(path, "wb").write(raw)
    offs, p = [], raw.find(b"AGFC")
    while p >= 0:
        offs.append(p)
        cs = int.from_bytes(raw[p + 12:p + 16], "little")
        p = raw.find(b"AGFC", p + 16 + cs)
    assert len(offs) == n_enc >= 56
    straight = list(range(n_enc))
    jumps = list(range(0, 11)) + [4, 5, 6, 7, 8, 9] + [40, 41, 42] + [12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23] + [3] + list(range(44, n_enc)) + [0, 1, 2, 3]
    for plan in (straight, jumps):
        base, st = _play_visit(lib, libc, path, raw, offs, plan, 1)
        assert st == [0, 0, 0]
        for depth in (8, 5):
            got, st = _play_visit(lib, libc, path, raw, offs, plan, depth)
            assert got == base, (comp, depth, [i for i, (x, y) in enumerate(zip(got, base)) if x != y][:5])
            assert st[0] > len(plan) // 3 and st[1] >= 2, st
            if plan is jumps:
                assert st[2] >= 3, st
