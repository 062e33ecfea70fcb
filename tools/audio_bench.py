#!/usr/bin/env python
"""Device time of the audio chunk codec kernels against the HBM roofline.

    python tools/audio_bench.py [million_samples]

Kernel time comes from the library's own CUDA events around each launch (agmvb_profile, class "audio");
algorithmic bytes: compress 2 B in + 1 B out per sample, expand 1 B in + 2 B out. Prints one JSON line.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libagmv_b200  # noqa: E402


def audio_ms(ctx):
    return ctx.profile_read()["audio"]


def main():
    n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 256_000_000
    peak = 6554.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ctx = libagmv_b200.Context(0)
    rng = np.random.default_rng(1)
    pcm = rng.integers(0, 65536, n, dtype=np.uint16)
    ctx.audio_compress(pcm[:1 << 20])  # warm-up, buffers
    out = {}
    for name, fn, arg in (("compress16", ctx.audio_compress, pcm), ("expand16", ctx.audio_expand, None)):
        if arg is None:
            arg = at
        fn(arg)  # sizes the buffers
        ctx.profile(1)
        for _ in range(3):
            res = fn(arg)
        cnt, ms = audio_ms(ctx)
        ctx.profile(0)
        if name == "compress16":
            at = res
        gbs = 3.0 * n * cnt / (ms * 1e-3) / 1e9
        out[name] = dict(launches=int(cnt), ms_per_launch=ms / cnt, algorithmic_gbs=gbs, frac_of_peak=gbs / peak)
    print(json.dumps(dict(samples=n, peak_gbs=peak, **out)))
    ctx.close()


if __name__ == "__main__":
    main()
