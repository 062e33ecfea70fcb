#!/usr/bin/env python
"""Frames per second through the player loop (AGMV_FindNextFrameChunk + AGMV_DecodeFrameChunk, i.e. AGMV_PlayAGMV,
src/agmv_playback.c:102-115) of the drop-in, with the frame-ahead queue off (AGMV_B200_AHEAD=1) and on.

    python tools/play_bench.py [w=1920] [h=1080] [n_source_frames=200] [depths=1,4,8,16]
"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libagmv_b200  # noqa: E402
from test_dropin import AGMV, _dropin, _libc  # noqa: E402  (ctypes view of the reference's handle)


def main():
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    depths = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "1,4,8,16").split(",")]
    ctx = libagmv_b200.Context(0)
    import torch
    dev = torch.empty((n, h, w), dtype=torch.int32, device="cuda:0")
    ctx.synth_frames(dev.data_ptr(), w, h, 0, n, 1234)
    data, n_enc = ctx.encode_sequence(None, n - 1, 24, 2, 2, 1, device_ptr=dev.data_ptr(), shape=(n, h, w))   # OPT_III, HIGH, LZSS
    ctx.close()
    raw = data.tobytes()
    lib, libc = _dropin(), _libc()
    res = {}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "p.agmv")
        open(path, "wb").write(raw)
        for depth in depths:
            os.environ["AGMV_B200_AHEAD"] = str(depth)
            f = libc.fopen(path.encode(), b"rb")
            hd = lib.CreateAGMV(0, w, h, 24)
            assert lib.AGMV_DecodeHeader(f, hd) == 0
            a = hd.contents
            a.frame.contents.width = a.iframe.contents.width = w
            a.frame.contents.height = a.iframe.contents.height = h
            a.frame_count = 0
            t0 = time.perf_counter()
            for k in range(n_enc):
                libc.fseek(f, raw.find(b"AGFC", libc.ftell(f)), 0)
                assert lib.AGMV_DecodeFrameChunk(f, hd) == 0, lib.AGMV_B200_LastError()
            dt = time.perf_counter() - t0
            libc.fclose(f)
            lib.DestroyAGMV(hd)
            res[f"ahead_{depth}"] = round(n_enc / dt, 1)
    print(json.dumps({"tool": "play_bench", "w": w, "h": h, "frames": n_enc, "frames_per_s": res}))


if __name__ == "__main__":
    main()
