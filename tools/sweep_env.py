#!/usr/bin/env python
"""Run bench.py once per value of an environment variable and print the K4 kernel classes of each run (tuning aid).

    python tools/sweep_env.py AGMVB_LZ_TARGET 8 16 32 -- --frames 1000 --steps 2 --warmup 2
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    var = sys.argv[1]
    sep = sys.argv.index("--") if "--" in sys.argv else len(sys.argv)
    vals, extra = sys.argv[2:sep], sys.argv[sep + 1:]
    for v in vals:
        env = dict(os.environ)
        if v != "default":
            env[var] = v
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-e2e", "--no-cpu-baseline"] + extra, env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            k = d["detail"]["kernel_ms_per_step"]
            print(var, v, "ms/step", round(d["ms_per_step"], 1), "enc fps", round(d["detail"]["encode_source_fps"]), "dec fps", round(d["detail"]["decode_fps"]),
                  {x: round(y, 1) for x, y in k.items() if y >= 1.0}, "launches", d["gpu_launches"], flush=True)
        except Exception as e:
            print(var, v, "failed", e, r.stderr[-500:], flush=True)


if __name__ == "__main__":
    main()
