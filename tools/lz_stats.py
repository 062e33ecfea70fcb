#!/usr/bin/env python
"""ANALYSIS AID (uses the oracle; not product, not a test): dumps the pre-LZSS block bitstreams of the bench workload's
first encoded frames and prints the statistics the K4 design is sized by (match-length histogram, level-L group sizes,
in-window predecessor counts).

    python tools/lz_stats.py [n_source_frames=16] [w=1920] [h=1080]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import ctypes as C

from agmv_testlib import OPT, QUALITY, oracle, ptr, synth_frames, _u8p, _u16p, _u32p, _u64p

W = 65535


def bitstreams(n_src, w, h, opt="III", quality="HIGH"):
    lib = oracle()
    fr = synth_frames(w, h, n_src, seed=1234)
    q = QUALITY[quality]
    hist = np.zeros(lib.orc_max_clr(q) + 1, dtype=np.uint64)
    for k in range(n_src):
        lib.orc_histogram_add(ptr(hist, _u64p), ptr(fr[k], _u32p), w * h, q)
    pal0 = np.zeros(256, dtype=np.uint32)
    pal1 = np.zeros(256, dtype=np.uint32)
    lib.orc_build_palette(ptr(hist, _u64p), q, OPT[opt], ptr(pal0, _u32p), ptr(pal1, _u32p))
    out = []
    ient = None
    enc = []
    i = 0
    while True:  # LIGHT schedule
        a = fr[i]
        b = np.empty_like(a)
        lib.orc_interp(ptr(b, _u32p), ptr(fr[i + 1], _u32p), ptr(fr[i + 2], _u32p), w * h)
        enc += [a, b, fr[i + 3]]
        i += 4
        if i + 4 >= n_src:
            break
    for k, px in enumerate(enc):
        ent = np.zeros(w * h, dtype=np.uint16)
        px = np.ascontiguousarray(px)
        lib.orc_quantize_frame(ptr(px, _u32p), w * h, ptr(pal0, _u32p), ptr(pal1, _u32p), 1, ptr(ent, _u16p))
        isi = k % 4 == 0
        buf = np.zeros(w * h * 3, dtype=np.uint8)
        n = lib.orc_assemble(ptr(ent, _u16p), ptr(ient if ient is not None else ent, _u16p), w, h, int(isi), 1, ptr(pal0, _u32p), ptr(pal1, _u32p), ptr(buf, _u8p))
        if isi:
            ient = ent
        out.append(buf[:n].copy())
    return out


def exact_matches(d):
    """bestlen M[p] (0 or 3..15) by level refinement in numpy."""
    n = len(d)
    pad = np.concatenate([d, np.zeros(16, dtype=np.uint8)]).astype(np.uint64)
    pos = np.arange(n, dtype=np.int64)
    ids = np.zeros(n, dtype=np.int64)
    M = np.zeros(n, dtype=np.int64)
    stats = {}
    for L in range(1, 16):
        key = ids * 256 + pad[L - 1:L - 1 + n].astype(np.int64)
        _, ids = np.unique(key, return_inverse=True)
        order = np.lexsort((pos, ids))
        sid = ids[order]
        sp = pos[order]
        same = np.zeros(n, dtype=bool)
        same[1:] = (sid[1:] == sid[:-1]) & (sp[1:] - sp[:-1] <= W)
        ex = np.zeros(n, dtype=bool)
        ex[sp] = same
        ex &= (pos + L <= n)
        if L >= 3:
            M[ex] = L
        cnt = np.bincount(ids)
        stats[L] = (cnt, ids.copy(), ex.copy())
    return M, stats


def main():
    n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
    bss = bitstreams(n_src, w, h)
    os.makedirs("/tmp/lzbs", exist_ok=True)
    for k, d in enumerate(bss):
        d.tofile(f"/tmp/lzbs/f{k}.bin")
    for k in (0, 1, 2):
        d = bss[k]
        n = len(d)
        M, st = exact_matches(d)
        print(f"--- frame {k} ({'I' if k % 4 == 0 else 'P'}) usize {n}")
        print("  M hist:", np.bincount(M, minlength=16).tolist())
        # greedy parse
        i = 0
        vis = []
        while i < n:
            vis.append(i)
            i += M[i] if M[i] >= 3 else 1
        vis = np.array(vis)
        print(f"  tokens {len(vis)}  literals {(M[vis] < 3).sum()}  visited M hist:", np.bincount(M[vis], minlength=16).tolist())
        for L in (3, 4, 5, 6, 8, 10, 12, 15):
            cnt, ids, ex = st[L]
            sz = cnt[ids]  # group size per position
            bins = [1, 2, 9, 33, 129, 513, 2049, 8193, 32769, 1 << 30]
            hh = np.histogram(sz, bins=bins)[0]
            print(f"  L={L}: groups {len(cnt)}  exists {ex.sum()}  positions by group size {bins[:-1]}: {hh.tolist()}")
        # byte histogram of top 3-grams
        cnt, ids, ex = st[3]
        top = np.argsort(-cnt)[:12]
        for g in top:
            p = np.flatnonzero(ids == g)[0]
            print(f"    3-gram {bytes(d[p:p + 3]).hex()} x {cnt[g]}")


if __name__ == "__main__":
    main()
