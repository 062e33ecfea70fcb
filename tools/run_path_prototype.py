#!/usr/bin/env python
"""Groundwork for the next K4 step: an exact, sort-free matcher for the level-3 groups whose key
is three equal bytes - the long-run groups that dominate the "large" path - validated here on the CPU against the
reference's brute-force search (AGMV_LZSS, src/agmv_encode.c:106-177).

For a position y with data[y] == data[y+1] == data[y+2] == b let r(y) be the length of the run of b starting at y.
Every candidate x (earlier, within the window, same three bytes) has

    LCP(x, y) = min(r(x), r(y))                         if r(x) != r(y)
              = r + LCP(data[x+r:], data[y+r:])          if r(x) == r(y) == r

so the longest match of y needs only (1) range queries over run lengths inside the window - "largest r", "earliest
position with r >= v" - which a table of runs answers without looking at the members one by one (the members of run
[s, e) with r >= v are exactly the positions s .. e-v), and (2) for y with r(y) < 15, a comparison of the bytes after the
run against the candidates that have the SAME r (one per earlier run): on the GPU those are small groups again and can go
through the existing kernels.

    best(y) = min(cap, max over candidates)          cap = min(15, bytes left)
    r(y) >= 15:  best = min(15, largest r in the window);                earliest x with min(r(x), 15) == best
    r(y) = r < 15:
        B = best over same-r candidates of min(r + extension, cap)         (0 if there is none)
        A = r if some candidate has r(x) > r, else the largest r(x) below r (0 if there is none)
            [so max(A, r if a same-r candidate exists) = min(r, largest r(x) in the window): the run table alone gives
             the answer unless a same-r candidate extends past the run, B > r]
        best = max(A, B); if best > r the start is the earliest same-r candidate reaching it, otherwise the earliest
        candidate with r(x) >= best - same-r candidates included, they all match at least r bytes.

Run:  python tools/run_path_prototype.py        (prints the number of positions checked; raises on the first mismatch)
"""
import bisect
import random

MAXLEN, MINLEN = 15, 3


def brute(data, window):
    """The reference's search at every position: (length, start) of the first longest match, length >= 3, else None."""
    n = len(data)
    out = [None] * n
    for i in range(n):
        mx = min(MAXLEN, n - i)
        best, bstart = 0, 0
        for s in range(max(0, i - window), i):
            j = 0
            while j < mx and data[s + j] == data[i + j]:
                j += 1
            if j > best:
                best, bstart = j, s
        if best >= MINLEN:
            out[i] = (best, bstart)
    return out


def run_aware(data, window):
    """Matches of the positions that start three equal bytes, from the run table; None elsewhere (other groups)."""
    n = len(data)
    out = [None] * n
    # runs of length >= 3 per byte value: (start, end)
    runs = {}
    i = 0
    while i < n:
        j = i
        while j < n and data[j] == data[i]:
            j += 1
        if j - i >= 3:
            runs.setdefault(data[i], []).append((i, j))
        i = j
    for b, rl in runs.items():
        starts = [s for s, e in rl]

        def members(lo, hi):
            """(run index, first member, last member) of the runs with members (positions s..e-3) inside [lo, hi)."""
            k = max(0, bisect.bisect_right(starts, lo) - 1)
            res = []
            while k < len(rl) and rl[k][0] < hi:
                s, e = rl[k]
                a, z = max(s, lo), min(e - 3, hi - 1)
                if a <= z:
                    res.append((k, a, z))
                k += 1
            return res

        def largest_r(lo, hi):
            return max((rl[k][1] - a for k, a, z in members(lo, hi)), default=0)   # r is largest at the first member

        def earliest_with_r_at_least(lo, hi, v):
            for k, a, z in members(lo, hi):
                if rl[k][1] - a >= v:          # the first member inside the range has the largest r of this run
                    return a
            return None

        def same_r_candidates(lo, hi, r):
            res = []
            for k, a, z in members(lo, hi):
                x = rl[k][1] - r
                if a <= x <= z:
                    res.append(x)
            return res

        for s, e in rl:
            for y in range(s, e - 2):
                r = e - y
                cap = min(MAXLEN, n - y)
                lo, hi = max(0, y - window), y
                if r >= MAXLEN:
                    m = min(MAXLEN, largest_r(lo, hi))
                    if m >= MINLEN:
                        out[y] = (min(m, cap), earliest_with_r_at_least(lo, hi, m))
                    continue
                # r < 15
                B, bx = 0, None
                for x in same_r_candidates(lo, hi, r):       # ascending: ">" keeps the earliest
                    j = r
                    while j < cap and data[x + j] == data[y + j]:
                        j += 1
                    j = min(j, cap)
                    if j > B:
                        B, bx = j, x
                # candidates with r(x) != r: per run the members inside the range carry every r from e-z up to e-a
                A = 0
                for k, a, z in members(lo, hi):
                    top = rl[k][1] - a
                    if top > r:
                        A = max(A, r)
                    elif top == r:
                        if z > a:
                            A = max(A, r - 1)
                    else:
                        A = max(A, top)
                best = max(A, B)
                if best < MINLEN:
                    continue
                if best > r:
                    out[y] = (best, bx)
                else:
                    out[y] = (min(best, cap), earliest_with_r_at_least(lo, hi, best))
    return out


def check(data, window):
    want = brute(data, window)
    got = run_aware(data, window)
    checked = 0
    for i in range(len(data)):
        if i + 2 < len(data) and data[i] == data[i + 1] == data[i + 2]:
            assert got[i] == want[i], (i, got[i], want[i], bytes(data[max(0, i - 20):i + 20]))
            checked += 1
    return checked


def main():
    rng = random.Random(5)
    total = 0
    for trial in range(300):
        parts = []
        for _ in range(rng.randint(5, 60)):
            kind = rng.random()
            if kind < 0.55:
                parts.append(bytes([rng.choice([0x5E, 0x5E, 0x4E, 7])]) * rng.randint(1, 40))
            elif kind < 0.8:
                parts.append(bytes(rng.choice([0x5E, 0x4E, 7, 9]) for _ in range(rng.randint(1, 6))))
            else:
                parts.append(bytes(rng.randrange(256) for _ in range(rng.randint(1, 8))))
        data = b"".join(parts)
        total += check(data, rng.choice([20, 60, 150, 65535]))
    print(f"run-aware matcher == brute force on {total} run positions")


if __name__ == "__main__":
    main()
