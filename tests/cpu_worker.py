"""One CPU process of the host baseline: encode one synthetic 1080p clip (AGMV_OPT_III, AGMV_HIGH_QUALITY, LZSS)
and decode it, with either the oracle port (`port`) or the unmodified reference binaries (`ref`).
Used by bench.py's cpu_baseline / --impl reference legs only."""
import sys
import time

from agmv_testlib import LZSS, OPT, QUALITY, oracle_decode, oracle_encode, ref_decode_raw, ref_encode, synth_frames


def main():
    n_src, seed, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    w, h = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (1920, 1080)
    frames = synth_frames(w, h, n_src, seed=seed)
    t0 = time.perf_counter()
    if mode == "ref":
        data = ref_encode(frames, n_src - 1, 24, OPT["III"], QUALITY["HIGH"], LZSS)
        t1 = time.perf_counter()
        rc, dec = ref_decode_raw(data)
    else:
        data = oracle_encode(frames, n_src - 1, 24, OPT["III"], QUALITY["HIGH"], LZSS)
        t1 = time.perf_counter()
        rc, dec = oracle_decode(data)
    t2 = time.perf_counter()
    assert rc == 0
    print(f"encode_s {t1 - t0:.3f} decode_s {t2 - t1:.3f} bytes {len(data)} frames {dec.shape[0]}")


if __name__ == "__main__":
    main()
