#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (oracle/_ref, built by oracle/Makefile from /root/reference).

Only runs where /root/reference exists (the build container). The outputs -
small .agmv streams plus golden.json with sizes and sha256 digests - are
committed so that the CPU test-suite and the GPU box (which has no
/root/reference) can check the oracle and the CUDA path against the reference.

    python tests/golden/make_golden.py [--skip-high] [--only NAME]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from agmv_testlib import (GOLDEN_DIR, LZSS, OPT, QUALITY, REF_DIR, ref_decode_raw, ref_encode, scene_cut_frames, sha256,  # noqa: E402
                          synth_frames)

# name, w, h, n_frames, create_n, fps, opt, quality, keep_file
ENCODE_CASES = [
    ("syn64_III_LOW", 64, 64, 12, 11, 24, "III", "LOW", True),
    ("syn64_I_MID", 64, 64, 12, 11, 24, "I", "MID", True),
    ("syn64_II_LOW", 64, 64, 12, 11, 24, "II", "LOW", True),
    ("syn64_ANIM_LOW", 64, 64, 20, 19, 24, "ANIM", "LOW", True),
    ("syn96x80_III_LOW", 96, 80, 24, 23, 30, "III", "LOW", True),
    ("gba240_GBA_I_LOW", 240, 160, 40, 39, 16, "GBA_I", "LOW", True),
    ("nds240_NDS_LOW", 240, 160, 24, 23, 16, "NDS", "LOW", True),
    ("c1_320x240_I_LOW", 320, 240, 212, 212, 24, "I", "LOW", False),
    ("c2_gba_full_GBA_I_LOW", 240, 160, 212, 211, 16, "GBA_I", "LOW", False),
    ("syn64_III_HIGH", 64, 64, 12, 11, 24, "III", "HIGH", True),
]

# "next" row N1: the other two sequence encoders, on frames with scene cuts so that AGMV_EncodeVideo's similarity gate
# takes both branches. name, mode, w, h, n_frames, create_n, fps, opt, quality
MODE_CASES = [
    ("video64_III_LOW", "video", 64, 64, 24, 23, 24, "III", "LOW"),
    ("video64_I_LOW", "video", 64, 64, 24, 23, 24, "I", "LOW"),
    ("video64_II_MID", "video", 64, 64, 24, 23, 24, "II", "MID"),
    ("video240_GBA_I_LOW", "video", 240, 160, 24, 23, 16, "GBA_I", "LOW"),
    ("full64_III_LOW", "full", 64, 64, 24, 23, 24, "III", "LOW"),
    ("full64_ANIM_LOW", "full", 64, 64, 10, 10, 24, "ANIM", "LOW"),
]


# "next" row N2: the LZ77 entropy coder (stream versions 3 = dual palette, 4 = single palette), all three sequence encoders
# name, mode, w, h, n_frames, create_n, fps, opt, quality
LZ77_CASES = [
    ("lz77_64_III_LOW", "agmv", 64, 64, 12, 11, 24, "III", "LOW"),
    ("lz77_64_II_LOW", "agmv", 64, 64, 12, 11, 24, "II", "LOW"),
    ("lz77_96x80_I_MID", "agmv", 96, 80, 24, 23, 30, "I", "MID"),
    ("lz77_gba240_GBA_I_LOW", "agmv", 240, 160, 24, 23, 16, "GBA_I", "LOW"),
    ("lz77_video64_III_LOW", "video", 64, 64, 24, 23, 24, "III", "LOW"),
    ("lz77_full64_ANIM_LOW", "full", 64, 64, 10, 10, 24, "ANIM", "LOW"),
    ("lz77_320x240_III_LOW", "agmv", 320, 240, 20, 19, 24, "III", "LOW"),
]


# "next" row N3: seeking. stream (golden file), visiting order of the frames
SEEK_CASES = [
    ("gba240_GBA_I_LOW.agmv", [0, 1, 8, 9, 10, 4, 5, 6, 7, 3, 16, 17, 0, 1]),
    ("syn96x80_III_LOW.agmv", [0, 1, 2, 3, 12, 13, 14, 5, 6, 8, 9, 2]),
    ("lz77_320x240_III_LOW.agmv", [4, 5, 6, 0, 1, 9, 10, 11, 8]),
    ("syn64_II_LOW.agmv", [3, 4, 0, 5]),
]


# AGMV_EncodeAGMV over a frame range that does not start at 1 (start_frame / end_frame arguments): name -> first, last source frame
RANGE_CASES = {"start3_64_III_LOW": (3, 16, "III", "LOW", 13), "start2_64_I_LOW": (2, 15, "I", "LOW", 13)}


# "next" row N4: streams with an audio track. name, mode, w, h, n_frames, create_n, fps, opt, quality, compression, bits, samples, rate, channels
AUDIO_CASES = [
    ("aud64_III_LOW", "agmv", 64, 64, 12, 11, 24, "III", "LOW", 1, 16, 20000, 8000, 1),
    ("aud64_I_LOW_pcm8", "agmv", 64, 64, 12, 11, 24, "I", "LOW", 1, 8, 24000, 8000, 2),
    ("aud_full64_ANIM_LOW", "full", 64, 64, 10, 10, 24, "ANIM", "LOW", 1, 16, 22050, 11025, 2),
    ("aud_lz77_64_III_LOW", "agmv", 64, 64, 12, 11, 24, "III", "LOW", 2, 16, 20000, 8000, 1),
    # every branch of the adjusted frame count the chunk size is divided by (src/agmv_encode.c:2296-2353)
    ("aud_gba240_GBA_I_LOW", "agmv", 240, 160, 24, 23, 16, "GBA_I", "LOW", 1, 16, 30000, 8000, 1),
    ("aud_gba240_GBA_III_LOW", "agmv", 240, 160, 24, 23, 16, "GBA_III", "LOW", 1, 16, 30011, 8000, 1),
    ("aud_nds240_NDS_LOW", "agmv", 240, 160, 24, 23, 16, "NDS", "LOW", 1, 8, 29999, 8000, 2),
    ("aud_lz77_64_II_LOW_pcm8", "agmv", 64, 64, 20, 19, 24, "II", "LOW", 2, 8, 21001, 8000, 1),
]


def make_audio_cases(gold):
    from agmv_testlib import audio_chunk_size, ref_audio, ref_audio_track, synth_pcm
    gold["audio"] = {}
    # known answers over the whole domain: every 16-bit sample through AGMV_CompressAudio, every code through AGMV_DecodeAudioChunk
    table = ref_audio("compress", np.arange(65536, dtype=np.uint16))
    table.tofile(os.path.join(GOLDEN_DIR, "audio_compress16.bin"))
    back = ref_audio("expand", np.arange(256, dtype=np.uint8))
    gold["audio"]["compress16_all"] = dict(file="audio_compress16.bin", sha256=sha256(table.tobytes()))
    gold["audio"]["expand16_all"] = [int(v) for v in back]
    gold["audio"]["streams"] = {}
    for name, mode, w, h, n, create_n, fps, opt, q, comp, bits, ns, rate, chn in AUDIO_CASES:
        frames = synth_frames(w, h, n, seed=1234)
        pcm = synth_pcm(ns, bits)
        data = ref_encode(frames, create_n, fps, OPT[opt], QUALITY[q], comp, mode=mode, audio=(pcm, rate, chn))
        # AGMV_EncodeFullAGMV sizes its chunks by end - start but writes end - start + 1 of them (src/agmv_encode.c:4024 vs :4041),
        # so its last chunk runs off the end of atsample into whatever the heap holds. Those bytes are not a function of the
        # input: they are zeroed in the fixture (and defined as zero in our implementation) and counted in "overread".
        n_enc = int.from_bytes(data[4:8], "little") if mode == "agmv" else n
        chunk = audio_chunk_size(ns, n, OPT[opt], mode)
        overread = max(0, chunk * n_enc - ns)
        if overread:
            assert overread <= chunk
            data = data[:-overread] + bytes(overread)
        rc, track = ref_audio_track(data)
        rc2, dec = ref_decode_raw(data)
        assert rc == 0 and rc2 == 0
        with open(os.path.join(GOLDEN_DIR, name + ".agmv"), "wb") as f:
            f.write(data)
        gold["audio"]["streams"][name] = dict(mode=mode, w=w, h=h, n=n, create_n=create_n, fps=fps, opt=opt, quality=q, compression=comp,
                                              bits=bits, samples=ns, rate=rate, channels=chn, pcm_sha256=sha256(pcm.tobytes()), overread=overread,
                                              size=len(data), sha256=sha256(data), file=name + ".agmv", track_samples=int(track.size),
                                              track_sha256=sha256(track.tobytes()),
                                              decoded_frame_sha256=[sha256(dec[k].tobytes()) for k in range(dec.shape[0])])
        print(f"audio {name}: {len(data)} B, track {track.size} samples", flush=True)


def make_range_cases(gold):
    """16 synthetic 64x64 frames written as f1..f16.bmp, the reference run on f<start>..f<end> (oracle/ref_encode.c)."""
    import subprocess
    import tempfile
    from agmv_testlib import write_bmps
    w, h, n = 64, 64, 16
    fr = synth_frames(w, h, n, seed=1234)
    gold["encode_ranges"] = {}
    for name, (start, end, opt, q, create_n) in RANGE_CASES.items():
        with tempfile.TemporaryDirectory() as td:
            write_bmps(fr, td, "f", 1)
            subprocess.run([os.path.join(REF_DIR, "ref_encode"), "o.agmv", ".", "f", str(start), str(end), str(w), str(h), "24", str(OPT[opt]),
                            str(QUALITY[q]), "1", str(create_n), "agmv"], cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            data = open(os.path.join(td, "o.agmv"), "rb").read()
        rc, dec = ref_decode_raw(data)
        assert rc == 0
        with open(os.path.join(GOLDEN_DIR, name + ".agmv"), "wb") as f:
            f.write(data)
        gold["encode_ranges"][name] = dict(w=w, h=h, n=n, seed=1234, start=start, end=end, opt=opt, quality=q, create_n=create_n, fps=24,
                                           size=len(data), sha256=sha256(data), file=name + ".agmv", decoded_shape=list(dec.shape))


def lzss_vectors():
    """Known-answer tests for the exported AGMV_LZSS (src/agmv_encode.c:106-177)."""
    rng = np.random.default_rng(7)
    vecs = {}
    vecs["random_4k"] = rng.integers(0, 256, 4096, dtype=np.uint8)
    vecs["four_symbols_20k"] = rng.integers(0, 4, 20000, dtype=np.uint8)
    vecs["all_equal_5k"] = np.full(5000, 0x5E, dtype=np.uint8)
    vecs["period5_3k"] = np.tile(np.array([1, 2, 3, 4, 5], dtype=np.uint8), 600)
    a = rng.integers(0, 256, 70000, dtype=np.uint8)
    a[66000:66040] = a[465:505]      # distance 65535: reachable
    a[67000:67040] = a[1464:1504]    # distance 65536: one past the window
    a[68000:68040] = a[2466:2506]    # distance 65534
    vecs["window_edges_70k"] = a
    vecs["empty"] = np.zeros(0, dtype=np.uint8)
    vecs["two_bytes"] = np.array([9, 9], dtype=np.uint8)
    vecs["tail_run"] = np.concatenate([rng.integers(0, 256, 100, dtype=np.uint8), np.full(17, 7, dtype=np.uint8)])
    return vecs


def ref_lzss(buf):
    """Call the reference's AGMV_LZSS + AGMV_FlushWriteBits through ctypes on a temp FILE*."""
    lib = C.CDLL(os.path.join(REF_DIR, "libagmv_ref.so"))
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]

    class BS(C.Structure):  # AGMV_BITSTREAM, include/agmv_defines.h:140-144 (u32 = unsigned long)
        _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_ulong), ("pos", C.c_ulong)]

    lib.AGMV_LZSS.restype = C.c_ulong
    lib.AGMV_LZSS.argtypes = [C.c_void_p, C.POINTER(BS)]
    lib.AGMV_FlushWriteBits.argtypes = [C.c_void_p]
    arr = np.concatenate([buf, np.zeros(32, dtype=np.uint8)])
    bs = BS(arr.ctypes.data_as(C.POINTER(C.c_uint8)), len(arr), len(buf))
    path = b"/tmp/_agmv_lzss_golden.bin"
    f = libc.fopen(path, b"wb")
    csize = lib.AGMV_LZSS(f, C.byref(bs))
    lib.AGMV_FlushWriteBits(f)
    libc.fclose(f)
    with open(path, "rb") as fh:
        out = fh.read()
    os.unlink(path)
    return int(csize), out


def ref_lz77(buf, stale=0xA7):
    """Call the reference's AGMV_LZ77 (src/agmv_encode.c:179-238) on a temp FILE*; data[pos] (read when the last match ends
    exactly at the end of the buffer) is set to `stale`."""
    lib = C.CDLL(os.path.join(REF_DIR, "libagmv_ref.so"))
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]

    class BS(C.Structure):
        _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_ulong), ("pos", C.c_ulong)]

    lib.AGMV_LZ77.restype = C.c_ulong
    lib.AGMV_LZ77.argtypes = [C.c_void_p, C.POINTER(BS)]
    arr = np.concatenate([buf, np.full(32, stale, dtype=np.uint8)])
    bs = BS(arr.ctypes.data_as(C.POINTER(C.c_uint8)), len(arr), len(buf))
    path = b"/tmp/_agmv_lz77_golden.bin"
    f = libc.fopen(path, b"wb")
    csize = lib.AGMV_LZ77(f, C.byref(bs))
    libc.fclose(f)
    with open(path, "rb") as fh:
        out = fh.read()
    os.unlink(path)
    return int(csize), out


def lz77_vectors():
    rng = np.random.default_rng(11)
    vecs = dict(lzss_vectors())
    vecs["all_equal_600"] = np.full(600, 0x5E, dtype=np.uint8)          # 255-byte matches, last one ends at the buffer end
    vecs["ends_in_match"] = np.concatenate([rng.integers(0, 256, 300, dtype=np.uint8)] * 2)   # second half = one long match chain
    vecs["period7_2k"] = np.tile(np.array([9, 8, 7, 6, 5, 4, 3], dtype=np.uint8), 300)
    vecs["one_byte"] = np.array([42], dtype=np.uint8)
    return vecs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-high", action="store_true")
    ap.add_argument("--only")
    args = ap.parse_args()
    jpath = os.path.join(GOLDEN_DIR, "golden.json")
    gold = json.load(open(jpath)) if os.path.exists(jpath) else {}
    gold.setdefault("encode", {})
    gold.setdefault("lzss", {})
    gold.setdefault("decode_fixture", {})

    for name, w, h, n, create_n, fps, opt, q, keep in ENCODE_CASES:
        if args.only and args.only != name:
            continue
        if args.skip_high and q == "HIGH":
            continue
        t0 = time.time()
        frames = synth_frames(w, h, n, seed=1234)
        timing = {}
        data = ref_encode(frames, create_n, fps, OPT[opt], QUALITY[q], LZSS, timing=timing)
        rc, dec = ref_decode_raw(data)
        assert rc == 0
        entry = dict(w=w, h=h, n=n, create_n=create_n, fps=fps, opt=opt, quality=q, seed=1234,
                     size=len(data), sha256=sha256(data), ref_encode_seconds=timing.get("seconds"),
                     decoded_shape=list(dec.shape), decoded_sha256=sha256(dec.tobytes()),
                     decoded_frame_sha256=[sha256(dec[k].tobytes()) for k in range(dec.shape[0])])
        if keep:
            with open(os.path.join(GOLDEN_DIR, name + ".agmv"), "wb") as f:
                f.write(data)
            entry["file"] = name + ".agmv"
        gold["encode"][name] = entry
        print(f"{name}: {len(data)} B, {dec.shape[0]} frames, {time.time() - t0:.1f}s", flush=True)
        json.dump(gold, open(jpath, "w"), indent=1, sort_keys=True)

    gold.setdefault("encode_modes", {})
    for name, mode, w, h, n, create_n, fps, opt, q in MODE_CASES:
        if args.only and args.only != name:
            continue
        if name in gold["encode_modes"] and not args.only:
            continue
        t0 = time.time()
        frames = scene_cut_frames(w, h, n)
        data = ref_encode(frames, create_n, fps, OPT[opt], QUALITY[q], LZSS, mode=mode)
        rc, dec = ref_decode_raw(data)
        assert rc == 0
        with open(os.path.join(GOLDEN_DIR, name + ".agmv"), "wb") as f:
            f.write(data)
        gold["encode_modes"][name] = dict(mode=mode, w=w, h=h, n=n, create_n=create_n, fps=fps, opt=opt, quality=q, size=len(data),
                                          sha256=sha256(data), file=name + ".agmv", frames_field=int.from_bytes(data[4:8], "little"),
                                          fps_field=int.from_bytes(data[18:22], "little"), decoded_shape=list(dec.shape),
                                          decoded_frame_sha256=[sha256(dec[k].tobytes()) for k in range(dec.shape[0])])
        print(f"{name}: {len(data)} B, {dec.shape[0]} frames, {time.time() - t0:.1f}s", flush=True)
        json.dump(gold, open(jpath, "w"), indent=1, sort_keys=True)

    gold.setdefault("encode_lz77", {})
    for name, mode, w, h, n, create_n, fps, opt, q in LZ77_CASES:
        if args.only and args.only != name:
            continue
        if name in gold["encode_lz77"] and not args.only:
            continue
        t0 = time.time()
        frames = scene_cut_frames(w, h, n) if mode != "agmv" else synth_frames(w, h, n, seed=1234)
        data = ref_encode(frames, create_n, fps, OPT[opt], QUALITY[q], 2, mode=mode)
        rc, dec = ref_decode_raw(data)
        assert rc == 0
        with open(os.path.join(GOLDEN_DIR, name + ".agmv"), "wb") as f:
            f.write(data)
        gold["encode_lz77"][name] = dict(mode=mode, w=w, h=h, n=n, create_n=create_n, fps=fps, opt=opt, quality=q, size=len(data),
                                         sha256=sha256(data), file=name + ".agmv", version=data[17], decoded_shape=list(dec.shape),
                                         decoded_frame_sha256=[sha256(dec[k].tobytes()) for k in range(dec.shape[0])])
        print(f"{name}: {len(data)} B, version {data[17]}, {dec.shape[0]} frames, {time.time() - t0:.1f}s", flush=True)
        json.dump(gold, open(jpath, "w"), indent=1, sort_keys=True)

    if not args.only or args.only == "audio":
        make_audio_cases(gold)
        json.dump(gold, open(jpath, "w"), indent=1, sort_keys=True)
    if not args.only:
        make_range_cases(gold)
        from agmv_testlib import ref_decode_seek
        gold["seek"] = {}
        for fn, plan in SEEK_CASES:
            data = open(os.path.join(GOLDEN_DIR, fn), "rb").read()
            rc, dec = ref_decode_seek(data, plan)
            assert rc == 0 and dec.shape[0] == len(plan)
            gold["seek"][fn] = dict(plan=plan, frame_sha256=[sha256(dec[k].tobytes()) for k in range(len(plan))])
            print(f"seek {fn}: {len(plan)} frames", flush=True)
        gold.setdefault("lz77", {})
        for name, buf in lz77_vectors().items():
            csize, out = ref_lz77(buf)
            gold["lz77"][name] = dict(n=len(buf), csize=csize, nbytes=len(out), sha256=sha256(out), input_sha256=sha256(buf.tobytes()), stale=0xA7)
            print(f"lz77 {name}: n={len(buf)} csize={csize} nbytes={len(out)}", flush=True)

        for name, buf in lzss_vectors().items():
            csize, out = ref_lzss(buf)
            gold["lzss"][name] = dict(n=len(buf), csize=csize, nbytes=len(out), sha256=sha256(out),
                                      input_sha256=sha256(buf.tobytes()))
            print(f"lzss {name}: n={len(buf)} csize={csize} nbytes={len(out)}", flush=True)

        # decode-only fixtures shipped with the reference (not copied: digests only)
        for rel in ["agmv_splash.agmv", "examples/simple_decoding/FOXLOGO.agmv", "agmv_spash.agmv"]:
            p = os.path.join("/root/reference", rel)
            data = open(p, "rb").read()
            rc, dec = ref_decode_raw(data)
            e = dict(rc=rc, input_sha256=sha256(data), size=len(data))
            if dec is not None and rc == 0:
                e.update(decoded_shape=list(dec.shape), decoded_sha256=sha256(dec.tobytes()))
            gold["decode_fixture"][rel] = e
            print(f"fixture {rel}: rc={rc}", flush=True)
    json.dump(gold, open(jpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
