#!/usr/bin/env python
"""Turn gpurun_out/*_<tag>.* (written by tools/ncu_capture.sh on the GPU box) into the tracked summaries under profiles/:
  profiles/<tag>_launches.csv     every kernel of one bench run: launches, total and mean device time, share
  profiles/<tag>_ncu_summary.txt  key ncu metrics of the --set full captures of the hot kernels
  profiles/traffic.json           dram bytes (read + write) per launch per kernel class, read by bench.py (roofline.traffic)
Run here (ncu is installed, no GPU needed): python tools/make_profiles.py r01
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
CLASS_OF = {"scatter": "rx_scatter", "group": "lz_group", "small": "lz_small", "tiny": "lz_tiny", "expand": "expand", "quantize": "quantize", "hist": "hist", "recon": "reconstruct",
            "pack": "lz_pack", "classify": "classify", "index": "index"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    # ---- launch list ----
    rows = list(csv.reader(open(os.path.join(OUT, f"launches_{tag}.csv"))))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[kn])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --frames 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline\n")
        f.write("# (cold-cache, serialised replays: compare SHARES, not absolutes; the run holds warm-up + timed + profiled + encode-only + decode-only passes)\n")
        f.write("kernel,launches,total_ms,mean_us,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1] / 1e6:.3f},{v[1] / v[0] / 1e3:.2f},{100 * v[1] / tot:.2f}\n")
    # ---- full captures ----
    traffic = {}
    with open(os.path.join(PROF, f"{tag}_ncu_summary.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one or two launches per hot kernel (tools/ncu_capture.sh)\n")
        for fn in sorted(os.listdir(OUT)):
            m = re.match(rf"prof_{tag}_(\w+)\.ncu-rep", fn)
            if not m:
                continue
            out = subprocess.run(["ncu", "-i", os.path.join(OUT, fn), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rr = list(csv.reader(io.StringIO(out)))
            if len(rr) < 3:
                continue
            h, units = rr[0], rr[1]
            per = []
            for r in rr[2:]:
                d, u = dict(zip(h, r)), dict(zip(h, units))
                f.write(f"== {m.group(1)}: {d.get('Kernel Name', '?')[:110]}\n")
                for k in KEYS:
                    if k in d:
                        f.write(f"   {k:80s} {d[k]:>16s} {u.get(k, '')}\n")
                rd = float(d["dram__bytes_read.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_read.sum"], 1)
                wr = float(d["dram__bytes_write.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_write.sum"], 1)
                per.append(rd + wr)
            if m.group(1) in CLASS_OF and per:
                traffic[CLASS_OF[m.group(1)]] = sum(per) / len(per)
    ratio = {}
    try:
        plain = json.load(open(os.path.join(OUT, f"plain_{tag}.json")))
        alg = plain["detail"]["alg_bytes_per_launch"]
        ratio = {k: traffic[k] / alg[k] for k in traffic if k in alg and alg[k] > 0}
    except Exception as e:
        print("no algorithmic bytes in the plain run:", e)
    json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full) of bench.py --frames 256; "
                       "traffic_over_algorithmic = that / the algorithmic bytes per launch of the same run",
               "dram_bytes_per_launch": traffic, "traffic_over_algorithmic": ratio},
              open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print(open(os.path.join(PROF, f"{tag}_launches.csv")).read()[:2500])


if __name__ == "__main__":
    main()
