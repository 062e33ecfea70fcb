#!/usr/bin/env python
"""profiles/traffic.json from a digest written by tools/ncu_capture.sh (gpurun_out/ncu_digest_<tag>.txt): DRAM bytes per captured
launch and, per kernel class, DRAM traffic over the algorithmic bytes bench.py charges the captured launches (DESIGN.md
section 5). bench.py multiplies that ratio by its own algorithmic bytes per launch for `roofline.traffic`.

    python tools/digest_to_traffic.py r02g
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the captured launches belong to the first encode batch of the bench configuration (1070 of the 1497 encoded frames,
# 517 M bitstream positions) and to the decode of all 1497 frames
BATCH1_FRAMES, ALL_FRAMES, BATCH1_POSITIONS, P = 1070, 1497, 0.517e9, 1920 * 1080


def main():
    tag = sys.argv[1]
    text = open(os.path.join(ROOT, "gpurun_out", f"ncu_digest_{tag}.txt")).read()
    plain = json.loads(open(os.path.join(ROOT, "gpurun_out", f"plain_{tag}.json")).read().strip().splitlines()[-1])
    alg = plain["detail"]["alg_bytes_per_launch"]
    launches = []
    for block in re.split(r"^== ", text, flags=re.M)[1:]:
        lines = block.split("\n")
        name = re.sub(r"\(.*", "", lines[0]).replace("void ", "").strip()
        vals = {}
        for ln in lines[1:]:
            m = re.match(r"\s+(.+?)\s{2,}([\d.]+)\s*(\S*)", ln)
            if m:
                vals[m.group(1).strip()] = float(m.group(2))
        if "time" not in vals:
            continue
        launches.append({"kernel": name, "time_ms": round(vals["time"], 3), "dram_read_GB": round(vals.get("dram read", 0.0), 3),
                         "dram_write_GB": round(vals.get("dram write", 0.0), 3),
                         "dram_GBps": round((vals.get("dram read", 0.0) + vals.get("dram write", 0.0)) / vals["time"] * 1e3, 1)})

    def total(prefix, first_only=False):
        sel = [x for x in launches if x["kernel"].startswith(prefix)]
        if first_only:
            sel = sel[:1]
        return sum(x["dram_read_GB"] + x["dram_write_GB"] for x in sel) * 1e9, len(sel)

    usize_all = alg["expand"] - plain["config"]["stream_bytes_per_gpu"] + 24 * ALL_FRAMES   # expand is charged csize + usize
    ratio = {}
    t, n = total("lzc_level_k")
    if n:
        ratio["lz_level"] = t / (n * BATCH1_POSITIONS)
    for cls, prefix, charged in (("lz_link", "lzc_hashlink_k", BATCH1_POSITIONS), ("lz_link3", "lzc_link3_k", BATCH1_POSITIONS),
                                 ("hist", "hist_vec4_k", alg["hist"]), ("expand", "expand_mrr_k", alg["expand"]),
                                 ("quantize", "quantize8_k", (4 * 4 / 3 + 2) * P * BATCH1_FRAMES),
                                 ("classify", "classify_k", (2 + 2 * 3 / 4) * P * BATCH1_FRAMES)):
        t, n = total(prefix, first_only=True)
        if n:
            ratio[cls] = t / charged
    t1, n1 = total("reconstruct_k", first_only=True)
    t2, n2 = total("recon_p_k", first_only=True)
    if n1:
        ratio["reconstruct"] = (t1 + t2) / (usize_all + 4.0 * P * ALL_FRAMES)
    out = {"source": f"gpurun_out/ncu_digest_{tag}.txt -> profiles/r02_ncu_summary.txt (ncu --set full at the bench configuration; first encode batch = "
                     "1070 frames, 517 M bitstream positions; decode = all 1497 frames)",
           "launches": launches, "traffic_over_algorithmic": {k: round(v, 3) for k, v in ratio.items()},
           "note": "traffic_over_algorithmic: dram__bytes_read.sum + dram__bytes_write.sum of the captured launch(es) over the algorithmic bytes "
                   "bench.py charges them (DESIGN.md section 5); lz_level = the twelve levels of the first batch together; reconstruct = "
                   "reconstruct_k + recon_p_k of one decode"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out["traffic_over_algorithmic"], indent=1))


if __name__ == "__main__":
    main()
