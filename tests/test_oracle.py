"""CPU tests: the C restatement (oracle/agmv_oracle.c) against the golden vectors
produced by the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from agmv_testlib import (GOLDEN_DIR, LZSS, OPT, QUALITY, have_ref, oracle_decode, oracle_encode, oracle_lzss, ref_decode_raw,
                          ref_encode, sha256, synth_frames)
from golden.make_golden import lzss_vectors

FAST_CASES = ["syn64_III_LOW", "syn64_I_MID", "syn64_II_LOW", "syn64_ANIM_LOW", "syn96x80_III_LOW", "gba240_GBA_I_LOW",
              "nds240_NDS_LOW", "syn64_III_HIGH"]


@pytest.mark.parametrize("name", FAST_CASES)
def test_oracle_encode_matches_reference_golden(golden, name):
    g = golden["encode"][name]
    frames = synth_frames(g["w"], g["h"], g["n"], seed=g["seed"])
    data = oracle_encode(frames, g["create_n"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZSS)
    assert len(data) == g["size"]
    assert sha256(data) == g["sha256"]
    with open(os.path.join(GOLDEN_DIR, g["file"]), "rb") as f:
        assert f.read() == data


@pytest.mark.parametrize("name", FAST_CASES)
def test_oracle_decode_matches_reference_golden(golden, name):
    g = golden["encode"][name]
    with open(os.path.join(GOLDEN_DIR, g["file"]), "rb") as f:
        data = f.read()
    rc, frames = oracle_decode(data)
    assert rc == 0
    assert list(frames.shape) == g["decoded_shape"]
    assert [sha256(frames[k].tobytes()) for k in range(frames.shape[0])] == g["decoded_frame_sha256"]


def test_oracle_c1_full_length(golden):
    """BASELINE config 1 (212 frames 320x240, OPT_I / LOW / LZSS): 104 encoded frames, byte-identical, then decode."""
    g = golden["encode"]["c1_320x240_I_LOW"]
    frames = synth_frames(g["w"], g["h"], g["n"], seed=g["seed"])
    data = oracle_encode(frames, g["create_n"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZSS)
    assert (len(data), sha256(data)) == (g["size"], g["sha256"])
    rc, dec = oracle_decode(data)
    assert rc == 0 and sha256(dec.tobytes()) == g["decoded_sha256"]


def test_oracle_lzss_known_answers(golden):
    for name, buf in lzss_vectors().items():
        g = golden["lzss"][name]
        assert sha256(buf.tobytes()) == g["input_sha256"], name
        csize, out, bits = oracle_lzss(buf)
        assert (csize, len(out), sha256(out)) == (g["csize"], g["nbytes"], g["sha256"]), name


def test_oracle_rejects_corrupt_header():
    # the shipped agmv_spash.agmv has bits_per_sample = 36904 and must be rejected with error 1
    hdr = bytearray(38 + 1536)
    hdr[0:4] = b"AGMV"
    hdr[17] = 1
    hdr[36:38] = (36904).to_bytes(2, "little")
    rc, _ = oracle_decode(bytes(hdr))
    assert rc == 1


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference fixtures live only in the build container")
def test_oracle_decodes_shipped_streams(golden):
    for rel, g in golden["decode_fixture"].items():
        data = open(os.path.join("/root/reference", rel), "rb").read()
        assert sha256(data) == g["input_sha256"]
        rc, frames = oracle_decode(data)
        assert rc == g["rc"], rel
        if rc == 0:
            assert sha256(frames.tobytes()) == g["decoded_sha256"], rel


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_oracle_matches_live_reference_random_profile():
    """A case that is NOT in the golden set, checked against the reference run live (fresh processes)."""
    frames = synth_frames(48, 32, 16, seed=99)
    ref = ref_encode(frames, 15, 20, OPT["III"], QUALITY["LOW"], LZSS)
    assert oracle_encode(frames, 15, 20, OPT["III"], QUALITY["LOW"], LZSS) == ref
    rc, rf = ref_decode_raw(ref)
    rc2, of = oracle_decode(ref)
    assert rc == rc2 == 0 and np.array_equal(rf, of)


MODE_CASES = ["video64_III_LOW", "video64_I_LOW", "video64_II_MID", "video240_GBA_I_LOW", "full64_III_LOW", "full64_ANIM_LOW"]


@pytest.mark.parametrize("name", MODE_CASES)
def test_oracle_other_encoders_match_reference_golden(golden, name):
    """AGMV_EncodeVideo (similarity-gated) and AGMV_EncodeFullAGMV restated: bytes and decoded frames vs the reference."""
    from agmv_testlib import oracle_encode_mode, scene_cut_frames
    g = golden["encode_modes"][name]
    frames = scene_cut_frames(g["w"], g["h"], g["n"])
    data = oracle_encode_mode(g["mode"], frames, g["create_n"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZSS)
    assert (len(data), sha256(data)) == (g["size"], g["sha256"])
    rc, dec = oracle_decode(data)
    assert rc == 0 and [sha256(dec[k].tobytes()) for k in range(dec.shape[0])] == g["decoded_frame_sha256"]


# ---- "next" row N2: the LZ77 entropy coder (stream versions 3/4) ----
LZ77_CASES = ["lz77_64_III_LOW", "lz77_64_II_LOW", "lz77_96x80_I_MID", "lz77_gba240_GBA_I_LOW", "lz77_video64_III_LOW",
              "lz77_full64_ANIM_LOW", "lz77_320x240_III_LOW"]


def test_oracle_lz77_known_answers(golden):
    """orc_lz77 against the reference's AGMV_LZ77 (src/agmv_encode.c:179-238), including the read one past the buffer."""
    from agmv_testlib import oracle_lz77
    from golden.make_golden import lz77_vectors
    for name, buf in lz77_vectors().items():
        g = golden["lz77"][name]
        assert sha256(buf.tobytes()) == g["input_sha256"], name
        csize, out = oracle_lz77(buf, g["stale"])
        assert (csize, len(out), sha256(out)) == (g["csize"], g["nbytes"], g["sha256"]), name


@pytest.mark.parametrize("name", LZ77_CASES)
def test_oracle_lz77_streams_match_reference_golden(golden, name):
    from agmv_testlib import LZ77, oracle_encode_mode, scene_cut_frames
    g = golden["encode_lz77"][name]
    if g["mode"] == "agmv":
        frames = synth_frames(g["w"], g["h"], g["n"], seed=1234)
        data = oracle_encode(frames, g["create_n"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZ77)
    else:
        frames = scene_cut_frames(g["w"], g["h"], g["n"])
        data = oracle_encode_mode(g["mode"], frames, g["create_n"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZ77)
    assert data[17] == g["version"]
    assert (len(data), sha256(data)) == (g["size"], g["sha256"])
    with open(os.path.join(GOLDEN_DIR, g["file"]), "rb") as f:
        assert f.read() == data
    rc, dec = oracle_decode(data)
    assert rc == 0 and list(dec.shape) == g["decoded_shape"]
    assert [sha256(dec[k].tobytes()) for k in range(dec.shape[0])] == g["decoded_frame_sha256"]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_seek_goldens_are_what_the_reference_produces(golden):
    """The seek vectors are re-derived from the unmodified reference (oracle/ref_decode.c seek mode), and
    visiting the frames in order reproduces the plain decode."""
    from agmv_testlib import ref_decode_seek
    for fn, g in golden["seek"].items():
        data = open(os.path.join(GOLDEN_DIR, fn), "rb").read()
        rc, dec = ref_decode_seek(data, g["plan"])
        assert rc == 0 and [sha256(dec[k].tobytes()) for k in range(dec.shape[0])] == g["frame_sha256"], fn
    g = golden["encode"]["syn96x80_III_LOW"]
    data = open(os.path.join(GOLDEN_DIR, g["file"]), "rb").read()
    rc, dec = ref_decode_seek(data, list(range(g["decoded_shape"][0])))
    assert rc == 0 and [sha256(dec[k].tobytes()) for k in range(dec.shape[0])] == g["decoded_frame_sha256"]


SMALL_SIZES = [(4, 4, 8), (8, 4, 12), (36, 20, 16), (132, 12, 9)]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_oracle_small_and_odd_frame_sizes_match_live_reference():
    """One block per frame, one block row, widths that are not multiples of 8 or 16: the oracle against the unmodified
    reference, bytes and decoded frames (the GPU test of the same name checks the CUDA path against the oracle)."""
    for w, h, n in (SMALL_SIZES[0], SMALL_SIZES[3]):          # the reference needs ~6 s per LOW-quality run (bubble sort)
        for opt, q in (("III", "LOW"), ("II", "LOW")):
            frames = synth_frames(w, h, n, seed=5)
            mine = oracle_encode(frames, n - 1, 24, OPT[opt], QUALITY[q], LZSS)
            ref = ref_encode(frames, n - 1, 24, OPT[opt], QUALITY[q], LZSS)
            assert mine == ref, (w, h, opt, q)
            rc, dec = oracle_decode(ref)
            rc2, dec2 = ref_decode_raw(ref)
            assert rc == rc2 == 0 and np.array_equal(dec, dec2), (w, h, opt, q)


@pytest.mark.parametrize("name", ["start3_64_III_LOW", "start2_64_I_LOW"])
def test_oracle_frame_range_not_starting_at_one(golden, name):
    """AGMV_EncodeAGMV(start_frame > 1): the reference on f<start>..f<end> equals the oracle on that slice of the frames."""
    g = golden["encode_ranges"][name]
    frames = synth_frames(g["w"], g["h"], g["n"], seed=g["seed"])[g["start"] - 1:g["end"]]
    data = oracle_encode(frames, g["create_n"], g["fps"], OPT[g["opt"]], QUALITY[g["quality"]], LZSS)
    assert (len(data), sha256(data)) == (g["size"], g["sha256"])
