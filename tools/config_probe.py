#!/usr/bin/env python
"""Single-GPU figures for the two BASELINE configurations bench.py does not time (they are not the headline workload):

  c4  3840x2160, AGMV_OPT_III / HIGH / LZSS: encode + decode of a device-resident prefix of the 4K sequence
      (what one rank does in the frame-range-sharded run), frames/s.
  c5  batched decode of S independent 1080p streams (each a config-3-style encode of 64 source frames -> 45 encoded
      frames, seeds 1000+s; D distinct streams, each opened S/D times - opens are independent decoder states):
      all streams advance together, one reconstruct launch per frame step for every stream (agmvb_dec_batch), frames/s;
      checked against the single-stream decoder (per-frame checksums), which tests/ hold to the oracle.

    python tools/config_probe.py [--c4-frames 256] [--c5-streams 512] [--c5-distinct 64] [--skip c4|c5]

Device time with CUDA events on the launching stream, inputs resident in HBM, 1 warm-up + 2 timed passes. One JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libagmv_b200  # noqa: E402

OPT_III, HIGH, LZSS = 3, 1, 1


def checksum_np(frames):
    """agmvb_dec_batch's per-frame checksum: sum over pixels of value * (2654435761 + 2 * pixel index), modulo 2^64."""
    out = []
    for f in frames:
        v = f.reshape(-1).astype(np.uint64)
        w = np.uint64(2654435761) + np.uint64(2) * np.arange(v.size, dtype=np.uint64)
        out.append(int((v * w).sum(dtype=np.uint64)))
    return np.array(out, dtype=np.uint64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4-frames", type=int, default=256)
    ap.add_argument("--c5-streams", type=int, default=512)
    ap.add_argument("--c5-distinct", type=int, default=64)
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = libagmv_b200.Context(0, stream.cuda_stream)
    res = {}

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    if "c4" not in args.skip:
        W, H, n = 3840, 2160, args.c4_frames
        frames = torch.empty((n, H, W), dtype=torch.int32, device=dev)
        ctx.synth_frames(frames.data_ptr(), W, H, 1, n, 1234)
        out = np.empty(2048 + n * (W * H // 2), np.uint8)
        enc = {}

        def encode():
            data, ne = ctx.encode_sequence(None, n - 1, 24, OPT_III, HIGH, LZSS, device_ptr=frames.data_ptr(), shape=(n, H, W), out=out)
            enc["data"], enc["n"] = data, ne

        enc_ms = timed(encode)
        data, ne = enc["data"], enc["n"]
        dec_out = torch.empty((ne, H, W), dtype=torch.int32, device=dev)

        def decode():
            sid, w, h, k = ctx.dec_open(data)
            ctx.dec_frames(sid, k, w, h, device_ptr=dec_out.data_ptr())
            ctx.dec_close(sid)

        dec_ms = timed(decode)
        res["c4"] = dict(workload=f"{n} source frames 3840x2160, OPT_III/HIGH/LZSS, device-resident, 1 GPU", encoded_frames=int(ne),
                         stream_bytes=int(len(data)), encode_ms=enc_ms, decode_ms=dec_ms, encode_source_fps=n / enc_ms * 1e3,
                         decode_fps=ne / dec_ms * 1e3, round_trip_source_fps=n / (enc_ms + dec_ms) * 1e3,
                         encode_compulsory_gbs=(8.0 * W * H * n + len(data)) / enc_ms / 1e6,
                         decode_compulsory_gbs=(4.0 * W * H * ne + len(data)) / dec_ms / 1e6)
        del frames, dec_out
        torch.cuda.empty_cache()

    if "c5" not in args.skip:
        W, H, S, D = 1920, 1080, args.c5_streams, args.c5_distinct
        src = torch.empty((64, H, W), dtype=torch.int32, device=dev)
        streams = []
        for s in range(D):
            ctx.synth_frames(src.data_ptr(), W, H, 1, 64, 1000 + s)
            data, ne = ctx.encode_sequence(None, 63, 24, OPT_III, HIGH, LZSS, device_ptr=src.data_ptr(), shape=(64, H, W))
            streams.append(np.array(data, copy=True))
        n_fr = int(ne)
        del src
        torch.cuda.empty_cache()
        state = {}

        def batched():
            sids = [ctx.dec_open(streams[s % D])[0] for s in range(S)]
            state["ck"] = ctx.dec_batch(sids, n_fr, None, checksums=True)
            for sid in sids:
                ctx.dec_close(sid)

        def batched_only():   # streams already open and uploaded: the reconstruction alone
            state["ck2"] = ctx.dec_batch(state["sids"], n_fr, None, checksums=True)

        all_ms = timed(batched)
        # decode-only: reopen, time one pass over already-uploaded streams
        state["sids"] = [ctx.dec_open(streams[s % D])[0] for s in range(S)]
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        batched_only()
        b.record(stream)
        torch.cuda.synchronize()
        only_ms = a.elapsed_time(b)
        for sid in state["sids"]:
            ctx.dec_close(sid)
        ck = state["ck"]
        # checks: replicas agree, batch == single-stream decoder
        for s in range(S):
            assert np.array_equal(ck[s], ck[s % D]) and np.array_equal(state["ck2"][s], ck[s % D]), f"stream {s} differs from its replica"
        ok_single = 0
        for s in range(0, D, max(1, D // 4)):
            single = ctx.decode_all(streams[s])
            assert np.array_equal(checksum_np(single), ck[s]), f"batched decode of stream {s} differs from the single-stream decoder"
            ok_single += 1
        total = S * n_fr
        sbytes = sum(len(streams[s % D]) for s in range(S))
        res["c5"] = dict(workload=f"{S} streams ({D} distinct, seeds 1000+s) x {n_fr} frames 1920x1080, OPT_III/HIGH/LZSS, batched decode, 1 GPU",
                         frames=total, open_upload_decode_ms=all_ms, decode_only_ms=only_ms, fps_with_open_and_upload=total / all_ms * 1e3,
                         fps_decode_only=total / only_ms * 1e3, decode_only_compulsory_gbs=(4.0 * W * H * total + sbytes) / only_ms / 1e6,
                         checked=f"replicas equal; {ok_single} streams vs the single-stream decoder (itself held to the oracle in tests/)")
    print(json.dumps(res))
    ctx.close()


if __name__ == "__main__":
    main()
