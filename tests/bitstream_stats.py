#!/usr/bin/env python
"""TEST INFRASTRUCTURE (analysis aid, uses the oracle): statistics of the LZSS layer on the bench workload's profile, to
size the K4 / D2 work per frame. Encodes a short 1080p OPT_III / HIGH / LZSS clip with the oracle port on the CPU and walks
every frame's token stream.

    python tests/bitstream_stats.py [n_source_frames=16]
"""
import json
import sys

import numpy as np

from agmv_testlib import LZSS, OPT, QUALITY, oracle_encode, synth_frames


def tokens(data, start, usize, csize):
    """Token list of one LZSS payload (src/agmv_decode.c:171-199): (is_literal, offset, length)."""
    bits = np.unpackbits(np.frombuffer(data[start:start + csize + 4], dtype=np.uint8), bitorder="little")
    p, nb, out, toks = 0, csize * 8, 0, []
    w16 = 1 << np.arange(16)
    w4 = 1 << np.arange(4)
    while p < nb and out < usize:
        if bits[p]:
            toks.append((1, 0, 1)); p += 9; out += 1
        else:
            off = int((bits[p + 1:p + 17] * w16).sum()); ln = int((bits[p + 17:p + 21] * w4).sum())
            toks.append((0, off, ln)); p += 21; out += ln
    return toks


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    frames = synth_frames(1920, 1080, n, seed=1234)
    data = oracle_encode(frames, n - 1, 24, OPT["III"], QUALITY["HIGH"], LZSS)
    o = 38 + 1536
    rows = []
    k = 0
    while o < len(data):
        assert data[o:o + 4] == b"AGFC"
        us, cs = int.from_bytes(data[o + 8:o + 12], "little"), int.from_bytes(data[o + 12:o + 16], "little")
        t = tokens(data, o + 16, us, cs)
        lit = sum(1 for a in t if a[0])
        m = [a for a in t if not a[0]]
        d = np.array([a[1] for a in m]) if m else np.zeros(0)
        ln = np.array([a[2] for a in m]) if m else np.zeros(0)
        rows.append(dict(frame=k, kind="I" if k % 4 == 0 else "P", usize=us, csize=cs, tokens=len(t), literals=lit, matches=len(m),
                         match_bytes=int(ln.sum()), mean_match_len=float(ln.mean()) if len(m) else 0.0,
                         dist_1=int((d == 1).sum()), dist_le_32=int((d <= 32).sum()), dist_le_4096=int((d <= 4096).sum()),
                         overlap=int((d < ln).sum())))
        o += 24 + cs + 8
        k += 1
    for r in rows:
        print(json.dumps(r))
    tot = {key: sum(r[key] for r in rows) for key in ("usize", "csize", "tokens", "literals", "matches", "match_bytes", "dist_1", "dist_le_32", "dist_le_4096", "overlap")}
    tot["frames"] = len(rows)
    tot["max_tokens_per_frame"] = max(r["tokens"] for r in rows)
    tot["mean_tokens_per_frame"] = tot["tokens"] / len(rows)
    print(json.dumps(dict(total=tot)))


if __name__ == "__main__":
    main()
