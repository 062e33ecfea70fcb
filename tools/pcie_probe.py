#!/usr/bin/env python
"""Host <-> device copy rates of the box: upload alone, download alone, both at once (two streams), from pinned memory,
for one and for several pairs of buffers in flight. The e2e leg of bench.py moves ~11 KB per source pixel-row through this
path, so these rates are its ceiling.

    python tools/pcie_probe.py [GB per buffer = 2]
"""
import sys
import time

import torch


def main():
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    n = int(gb * 1e9)
    dev = torch.device("cuda", 0)
    for pairs in (1, 2, 4):
        hs = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(pairs)]
        ho = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(pairs)]
        for h in hs + ho:
            h.fill_(1)
        ds = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(pairs)]
        do = [torch.ones(n, dtype=torch.uint8, device=dev) for _ in range(pairs)]
        up = [torch.cuda.Stream() for _ in range(pairs)]
        dn = [torch.cuda.Stream() for _ in range(pairs)]

        def run(do_up, do_dn, reps=3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                for k in range(pairs):
                    if do_up:
                        with torch.cuda.stream(up[k]):
                            ds[k].copy_(hs[k], non_blocking=True)
                    if do_dn:
                        with torch.cuda.stream(dn[k]):
                            ho[k].copy_(do[k], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            return reps * pairs * n * (int(do_up) + int(do_dn)) / dt / 1e9

        run(True, True, 1)
        print(f"{pairs} buffer pair(s) of {gb:.1f} GB: upload alone {run(True, False):.1f} GB/s, download alone {run(False, True):.1f} GB/s, "
              f"both at once {run(True, True):.1f} GB/s (sum of both directions)", flush=True)
        del hs, ho, ds, do


if __name__ == "__main__":
    main()
