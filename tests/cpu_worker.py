"""One CPU process of the host baseline: encode one synthetic clip (AGMV_OPT_III, LZSS) and decode it without export, with
either the oracle port (`port`) or the unmodified reference binaries built by oracle/Makefile (`ref`).
Used by bench.py's cpu_baseline / --impl reference legs only. Prints its own timers (frame synthesis and interpreter start-up
are outside them):

    encode_s E decode_s D palette_s P bytes B frames N

palette_s (port only) is the fixed cost inside encode_s: histogram over all frames + sort + greedy palette pick.
"""
import ctypes as C
import sys
import time

import numpy as np

from agmv_testlib import LZSS, OPT, QUALITY, oracle, oracle_decode, oracle_encode, ptr, ref_decode_raw, ref_encode, synth_frames


def main():
    n_src, seed, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    w, h = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (1920, 1080)
    quality = sys.argv[6] if len(sys.argv) > 6 else "HIGH"
    frames = synth_frames(w, h, n_src, seed=seed)
    pal_s = 0.0
    if mode == "ref":
        te, td = {}, {}
        data = ref_encode(frames, n_src - 1, 24, OPT["III"], QUALITY[quality], LZSS, timing=te)     # the binaries time AGMV_EncodeAGMV /
        rc, dec = ref_decode_raw(data, timing=td)                                                      # the per-frame decode loop themselves
        enc_s, dec_s = te["seconds"], td["seconds"]
    else:
        lib = oracle()
        q = QUALITY[quality]
        t0 = time.perf_counter()
        hist = np.zeros(lib.orc_max_clr(q) + 1, dtype=np.uint64)
        for k in range(n_src):
            lib.orc_histogram_add(ptr(hist, C.POINTER(C.c_uint64)), ptr(frames[k], C.POINTER(C.c_uint32)), w * h, q)
        p0, p1 = np.zeros(256, np.uint32), np.zeros(256, np.uint32)
        lib.orc_build_palette(ptr(hist, C.POINTER(C.c_uint64)), q, OPT["III"], ptr(p0, C.POINTER(C.c_uint32)), ptr(p1, C.POINTER(C.c_uint32)))
        pal_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        data = oracle_encode(frames, n_src - 1, 24, OPT["III"], q, LZSS)
        t1 = time.perf_counter()
        rc, dec = oracle_decode(data)
        t2 = time.perf_counter()
        enc_s, dec_s = t1 - t0, t2 - t1
    assert rc == 0
    print(f"encode_s {enc_s:.3f} decode_s {dec_s:.3f} palette_s {pal_s:.3f} bytes {len(data)} frames {dec.shape[0]}")


if __name__ == "__main__":
    main()
