"""The audio chunk codec: AGMV_CompressAudio / AGMV_EncodeAudioChunk / AGMV_DecodeAudioChunk.

CPU part: the oracle's restatement against the known answers of the unmodified reference - every 16-bit sample value
through AGMV_CompressAudio, every code through AGMV_DecodeAudioChunk, eight golden streams with a track (16- and 8-bit,
mono and stereo, AGMV_EncodeAGMV and AGMV_EncodeFullAGMV, both entropy coders, every profile family).
GPU part (-m gpu): the kernels behind the C-ABI and the drop-in API against the same vectors and the oracle.
"""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import libagmv_b200
from agmv_testlib import (GOLDEN_DIR, OPT, QUALITY, audio_chunk_size, have_ref, mux_audio, oracle_audio_compress16, oracle_audio_expand16,
                          oracle_encode, oracle_encode_mode, ref_audio, ref_audio_track, sha256, synth_frames, synth_pcm, write_bmps)

STREAMS = ["aud64_III_LOW", "aud64_I_LOW_pcm8", "aud_full64_ANIM_LOW", "aud_lz77_64_III_LOW",
           # one per branch of the adjusted frame count the chunk size divides by (GBA_I: / 2, GBA_III: * 0.75f, NDS / II: * 0.75)
           "aud_gba240_GBA_I_LOW", "aud_gba240_GBA_III_LOW", "aud_nds240_NDS_LOW", "aud_lz77_64_II_LOW_pcm8"]


def _table(golden):
    g = golden["audio"]["compress16_all"]
    t = np.fromfile(os.path.join(GOLDEN_DIR, g["file"]), dtype=np.uint8)
    assert t.size == 65536 and sha256(t.tobytes()) == g["sha256"]
    return t


def _track_of(stream):
    """The sample bytes of every audio chunk, walking the stream the way the reference's decode loop does
    (src/agmv_decode.c:583-587) - possible without a bit reader because every frame chunk carries its csize."""
    s = bytes(stream)
    o = 38 + 768 * (2 if s[17] in (1, 3) else 1)
    out = bytearray()
    while o < len(s):
        assert s[o:o + 4] == b"AGFC"
        o += 24 + int.from_bytes(s[o + 12:o + 16], "little")
        assert s[o:o + 4] == b"AGAC"
        n = int.from_bytes(s[o + 4:o + 8], "little")
        out += s[o + 8:o + 8 + n]
        o += 8 + n
    return np.frombuffer(bytes(out), dtype=np.uint8)


def _case(golden, name):
    g = golden["audio"]["streams"][name]
    frames = synth_frames(g["w"], g["h"], g["n"], seed=1234)
    pcm = synth_pcm(g["samples"], g["bits"])
    assert sha256(pcm.tobytes()) == g["pcm_sha256"], "synthetic track generator changed"
    with open(os.path.join(GOLDEN_DIR, g["file"]), "rb") as f:
        data = f.read()
    assert sha256(data) == g["sha256"]
    return g, frames, pcm, data


# ------------------------------------------------------------------------------------------------ CPU: oracle vs reference
def test_oracle_compress_every_sample_value(golden):
    assert np.array_equal(oracle_audio_compress16(np.arange(65536, dtype=np.uint16)), _table(golden))


def test_oracle_expand_every_code(golden):
    want = np.array(golden["audio"]["expand16_all"], dtype=np.uint16)
    assert np.array_equal(oracle_audio_expand16(np.arange(256, dtype=np.uint8)), want)
    # AGMV_SQR_TABLE / AGMV_SHIFT_TABLE (src/agmv_decode.c:21-89) are what the restatement says they are
    b = np.arange(256)
    assert np.array_equal(want, np.where(b % 2 == 0, b * b, b * 256).astype(np.uint16))


@pytest.mark.parametrize("name", STREAMS)
def test_oracle_streams_with_audio_match_reference_golden(golden, name):
    g, frames, pcm, data = _case(golden, name)
    args = (frames, g["create_n"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], g["compression"])
    base = oracle_encode(*args) if g["mode"] == "agmv" else oracle_encode_mode(g["mode"], *args)
    at = oracle_audio_compress16(pcm) if g["bits"] == 16 else pcm
    chunk = audio_chunk_size(pcm.size, g["n"], OPT[g["opt"]], g["mode"])
    assert mux_audio(base, at, chunk, pcm, g["rate"], g["channels"], g["mode"]) == data
    coded = _track_of(data)
    track = oracle_audio_expand16(coded) if g["bits"] == 16 else coded
    assert track.size == g["track_samples"] and sha256(track.tobytes()) == g["track_sha256"]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_audio_goldens_are_what_the_reference_produces(golden):
    rng = np.random.default_rng(3)
    pcm = rng.integers(0, 65536, 5000, dtype=np.uint16)
    assert np.array_equal(ref_audio("compress", pcm), oracle_audio_compress16(pcm))
    g, frames, pcm, data = _case(golden, "aud64_III_LOW")
    rc, track = ref_audio_track(data)
    assert rc == 0 and sha256(track.tobytes()) == g["track_sha256"]


# ------------------------------------------------------------------------------------------------ GPU: C-ABI
@pytest.mark.gpu
def test_compress_every_sample_value(ctx, golden):
    assert np.array_equal(ctx.audio_compress(np.arange(65536, dtype=np.uint16)), _table(golden))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 15, 16, 17, 31, 4097, 1_000_003, 3_000_017])  # the last one takes the shared-memory table kernel
def test_compress_and_expand_ragged_lengths(ctx, n):
    rng = np.random.default_rng(n)
    pcm = rng.integers(0, 65536, n, dtype=np.uint16)
    at = ctx.audio_compress(pcm)
    assert np.array_equal(at, oracle_audio_compress16(pcm))
    assert np.array_equal(ctx.audio_expand(at), oracle_audio_expand16(at))
    codes = rng.integers(0, 256, n, dtype=np.uint8)   # codes the encoder never emits decode like the reference's tables too
    assert np.array_equal(ctx.audio_expand(codes), oracle_audio_expand16(codes))
    p8 = rng.integers(0, 256, n, dtype=np.uint8)      # 8-bit tracks are stored as they are (src/agmv_encode.c:699-703)
    assert np.array_equal(ctx.audio_compress(p8), p8) and np.array_equal(ctx.audio_expand(p8, bits=8), p8)


@pytest.mark.gpu
def test_expand_every_code(ctx, golden):
    assert np.array_equal(ctx.audio_expand(np.arange(256, dtype=np.uint8)), np.array(golden["audio"]["expand16_all"], dtype=np.uint16))


@pytest.mark.gpu
@pytest.mark.parametrize("name", STREAMS)
def test_streams_with_audio_match_reference_golden(ctx, golden, name):
    g, frames, pcm, data = _case(golden, name)
    ctx.enc_set_audio(pcm, g["rate"], g["channels"], pcm.nbytes // (g["rate"] * g["channels"] * pcm.dtype.itemsize))
    args = (frames, g["create_n"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], g["compression"])
    mine, n_enc = ctx.encode_sequence(*args) if g["mode"] == "agmv" else ctx.encode_mode(g["mode"], *args)
    assert mine.tobytes() == data
    # the track is consumed with the handle: the next sequence has none
    again, _ = ctx.encode_sequence(*args) if g["mode"] == "agmv" else ctx.encode_mode(g["mode"], *args)
    base = oracle_encode(*args) if g["mode"] == "agmv" else oracle_encode_mode(g["mode"], *args)
    assert again.tobytes() == base
    # decode: pictures and track
    sid, w, h, n = ctx.dec_open(data)
    try:
        track = ctx.dec_audio(sid)
        dec = ctx.dec_frames(sid, n, w, h)
        assert np.array_equal(ctx.dec_audio(sid), track)   # independent of the frame cursor
    finally:
        ctx.dec_close(sid)
    assert track.size == g["track_samples"] and sha256(track.tobytes()) == g["track_sha256"]
    assert [sha256(dec[k].tobytes()) for k in range(n)] == g["decoded_frame_sha256"]


@pytest.mark.gpu
def test_stream_without_track_has_no_audio(ctx, golden):
    g = golden["encode"]["syn64_III_LOW"]
    with open(os.path.join(GOLDEN_DIR, g["file"]), "rb") as f:
        sid, w, h, n = ctx.dec_open(f.read())
    try:
        assert ctx.dec_audio(sid).size == 0
    finally:
        ctx.dec_close(sid)


@pytest.mark.gpu
def test_truncated_last_audio_chunk_reads_ff(ctx, golden):
    """AGIDL_ReadByte past the end of the file stores EOF in a u8: the missing codes are 0xFF -> 65280."""
    g, frames, pcm, data = _case(golden, "aud64_III_LOW")
    cut = data[:-100]
    sid, w, h, n = ctx.dec_open(cut)
    try:
        track = ctx.dec_audio(sid)
    finally:
        ctx.dec_close(sid)
    coded = np.concatenate([_track_of(data)[:-100], np.full(100, 0xFF, np.uint8)])
    assert np.array_equal(track, oracle_audio_expand16(coded))
    if have_ref():
        rc, rtrack = ref_audio_track(cut)
        assert rc == 0 and np.array_equal(rtrack, track)


@pytest.mark.gpu
def test_manual_chunk_interleave_matches_sequence_encoder(ctx, golden):
    """agmvb_enc_frames with agmvb_enc_set_audio_chunk (what the drop-in's AGMV_EncodeAGMV drives), in two calls."""
    g, frames, pcm, data = _case(golden, "aud64_III_LOW")
    chunk = audio_chunk_size(pcm.size, g["n"], OPT["III"])
    ctx.enc_begin(g["w"], g["h"], OPT["III"], QUALITY["LOW"], 1)
    fr = np.ascontiguousarray(frames)
    ctx.enc_histogram(fr.ctypes.data, g["n"], 0)
    ctx.enc_build_palette()
    ctx.enc_set_audio(pcm, g["rate"], g["channels"], 2)
    assert np.array_equal(ctx.enc_get_atsample(pcm.size), oracle_audio_compress16(pcm))
    ctx.enc_set_audio_chunk(chunk)
    body = b""
    sa, sb = [0, 1, 3, 4, 5, 7], [-1, 2, -1, -1, 6, -1]   # LIGHT schedule of 12 source frames: two groups
    for part in (slice(0, 3), slice(3, 6)):
        nbytes = ctx.enc_frames(fr.ctypes.data, g["n"], 0, sa[part], sb[part], first_frame_count=part.start)
        body += ctx.enc_fetch(nbytes, 3)[0].tobytes()
    ctx.enc_set_audio(None, 0, 0, 0)
    ctx.lib.agmvb_enc_set_audio_stub(ctx.h, 1)
    o = 38 + 1536
    assert body == data[o:]


# ------------------------------------------------------------------------------------------------ GPU: drop-in API
UL = C.c_ulong


class AudioChunk(C.Structure):
    _fields_ = [("fourcc", C.c_char * 4), ("size", UL), ("atsample", C.POINTER(C.c_uint8)), ("satsample", C.c_void_p)]


class AudioTrack(C.Structure):
    _fields_ = [("total_audio_duration", UL), ("start_point", UL), ("pcm", C.POINTER(C.c_uint16)), ("pcm8", C.POINTER(C.c_uint8))]


def _load_track(h, pcm, rate, channels):
    """What AGMV_WavToAudioTrack leaves in the handle (src/agmv_utils.c:1070-1083); the buffer is malloc'ed because
    AGMV_EncodeAGMV frees the handle, track included."""
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    buf = libc.malloc(pcm.nbytes)
    C.memmove(buf, pcm.ctypes.data, pcm.nbytes)
    hd = h.contents.header
    hd.total_audio_duration = pcm.nbytes // (rate * channels * pcm.dtype.itemsize)
    hd.sample_rate, hd.num_of_channels, hd.bits_per_sample, hd.audio_size = rate, channels, pcm.dtype.itemsize * 8, pcm.size
    tr = C.cast(h.contents.audio_track, C.POINTER(AudioTrack)).contents
    if pcm.dtype == np.uint16:
        tr.pcm = C.cast(buf, C.POINTER(C.c_uint16))
    else:
        tr.pcm8 = C.cast(buf, C.POINTER(C.c_uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("name", STREAMS)
def test_dropin_encode_with_audio_track(golden, name):
    """examples/simple_video_and_audio/simple_video_and_audio.c: CreateAGMV, a track in the handle, AGMV_EncodeAGMV."""
    from test_dropin import _dropin
    g, frames, pcm, data = _case(golden, name)
    lib = _dropin()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        write_bmps(frames, td, "f", 1)
        os.chdir(td)
        try:
            h = lib.CreateAGMV(g["create_n"], g["w"], g["h"], g["fps"])
            _load_track(h, pcm, g["rate"], g["channels"])
            fn = lib.AGMV_EncodeAGMV if g["mode"] == "agmv" else lib.AGMV_EncodeFullAGMV
            fn(h, b"o.agmv", b".", b"f", 1, 1, g["n"], g["w"], g["h"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], g["compression"])
            mine = open("o.agmv", "rb").read()
        finally:
            os.chdir(cwd)
    assert mine == data, lib.AGMV_B200_LastError()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["aud64_III_LOW", "aud64_I_LOW_pcm8"])
def test_dropin_decode_exports_wav(golden, name):
    """AGMV_DecodeAGMV(file, BMP, WAV) writes quick_export.wav (src/agmv_decode.c:589-594, src/agmv_utils.c:1407-1435)."""
    from test_dropin import _dropin
    g, frames, pcm, data = _case(golden, name)
    lib = _dropin()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        try:
            assert lib.AGMV_DecodeAGMV(os.path.join(GOLDEN_DIR, g["file"]).encode(), 1, 1) == 0
            wav = open("quick_export.wav", "rb").read()
            assert len([f for f in os.listdir(".") if f.endswith(".bmp")]) == len(g["decoded_frame_sha256"])
        finally:
            os.chdir(cwd)
    nbytes = pcm.nbytes
    assert wav[:4] == b"RIFF" and int.from_bytes(wav[4:8], "little") == nbytes and wav[8:16] == b"WAVEfmt "
    assert int.from_bytes(wav[22:24], "little") == g["channels"] and int.from_bytes(wav[24:28], "little") == g["rate"]
    assert int.from_bytes(wav[34:36], "little") == g["bits"] and wav[36:40] == b"data" and len(wav) == 44 + nbytes
    samples = np.frombuffer(wav[44:], dtype=np.uint16 if g["bits"] == 16 else np.uint8)
    assert sha256(samples[:g["track_samples"]].tobytes()) == g["track_sha256"]
    assert not samples[g["track_samples"]:].any()   # the reference leaves these uninitialised; zero here


@pytest.mark.gpu
def test_dropin_per_chunk_audio_api(golden):
    """AGMV_CompressAudio + AGMV_EncodeAudioChunk / AGMV_DecodeAudioChunk on a handle, as the reference's own loops call them."""
    from test_dropin import _dropin, _libc
    lib, libc = _dropin(), _libc()
    lib.AGMV_CompressAudio.argtypes = [C.c_void_p]
    lib.AGMV_EncodeAudioChunk.argtypes = [C.c_void_p, C.c_void_p]
    lib.AGMV_DecodeAudioChunk.argtypes = [C.c_void_p, C.c_void_p]
    lib.AGMV_DecodeAudioChunk.restype = C.c_int
    pcm = synth_pcm(6000, 16, seed=11)
    want = oracle_audio_compress16(pcm)
    h = lib.CreateAGMV(4, 16, 16, 8)
    _load_track(h, pcm, 1000, 1)
    ch = C.cast(h.contents.audio_chunk, C.POINTER(AudioChunk)).contents
    tr = C.cast(h.contents.audio_track, C.POINTER(AudioTrack)).contents
    at = (C.c_uint8 * pcm.size)()
    ch.atsample = C.cast(at, C.POINTER(C.c_uint8))
    lib.AGMV_CompressAudio(h)
    assert np.array_equal(np.frombuffer(at, dtype=np.uint8), want)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "a.bin").encode()
        f = libc.fopen(path, b"wb")
        ch.size, tr.start_point = 1500, 0
        for _ in range(4):
            lib.AGMV_EncodeAudioChunk(f, h)
        libc.fclose(f)
        raw = open(path, "rb").read()
        assert len(raw) == 4 * 1508 and all(raw[k * 1508:k * 1508 + 8] == b"AGAC" + (1500).to_bytes(4, "little") for k in range(4))
        assert b"".join(raw[k * 1508 + 8:(k + 1) * 1508] for k in range(4)) == want.tobytes()
        # decode the four chunks back into the track
        back = (C.c_uint16 * pcm.size)()
        keep = C.cast(tr.pcm, C.c_void_p).value   # the address, not a view of the field
        tr.pcm, tr.start_point = C.cast(back, C.POINTER(C.c_uint16)), 0
        f = libc.fopen(path, b"rb")
        for _ in range(4):
            assert lib.AGMV_DecodeAudioChunk(f, h) == 0
        assert lib.AGMV_DecodeAudioChunk(f, h) == 1   # no 'AGAC' at the end of the file
        libc.fclose(f)
        assert tr.start_point == 6000
        assert np.array_equal(np.frombuffer(back, dtype=np.uint16), oracle_audio_expand16(want))
        tr.pcm = C.cast(C.c_void_p(keep), C.POINTER(C.c_uint16))
    ch.atsample = None
    lib.DestroyAGMV(h)
