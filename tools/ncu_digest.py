#!/usr/bin/env python
"""Digest ncu reports where ncu is installed (the GPU box or the build container; no GPU needed): per captured launch the
metrics the roofline discussion uses, and optionally the source lines that collect the most stall samples.

    python tools/ncu_digest.py [--top N] report.ncu-rep [...]
"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads per instruction"),
        ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__occupancy_limit_registers", "occ limit regs"), ("launch__occupancy_limit_shared_mem", "occ limit smem"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
        ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall imc_miss"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
        ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall drain"),
        ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall tex_throttle"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
        ("smsp__inst_executed_op_local_ld.sum", "local loads"), ("smsp__inst_executed_op_local_st.sum", "local stores"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"), ("lts__t_sectors_data_ecc.sum", "L2 ECC fill sectors")]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    args = sys.argv[1:]
    top = 0
    if args and args[0] == "--top":
        top = int(args[1])
        args = args[2:]
    for path in args:
        rows = list(csv.reader(io.StringIO(run(["-i", path, "--page", "raw", "--csv"]))))
        if len(rows) < 3:
            print(path, "no data")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            print(f"== {d.get('Kernel Name', '?')[:100]}   [id {d.get('ID')}, {path.split('/')[-1]}]")
            for k, label in KEYS:
                if k in d and d[k] != "":
                    print(f"   {label:28s} {d[k]:>18s} {u.get(k, '')}")
        if top:
            src = list(csv.reader(io.StringIO(run(["-i", path, "--page", "source", "--csv"]))))
            # the source page lists one kernel after the other: "Kernel Name" rows separate them
            cur, blocks = None, []
            for r in src:
                if r and r[0] == "Kernel Name":
                    cur = {"name": r[1], "hdr": None, "rows": []}
                    blocks.append(cur)
                elif cur is not None and cur["hdr"] is None:
                    cur["hdr"] = r
                elif cur is not None and r:
                    cur["rows"].append(r)
            for b in blocks:
                ix = {k: i for i, k in enumerate(b["hdr"])}
                if "# Samples" not in ix:
                    continue
                tot = sum(int(r[ix["# Samples"]] or 0) for r in b["rows"]) or 1
                best = sorted(b["rows"], key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]
                print(f"-- top stall lines of {b['name'][:80]} ({tot} samples)")
                for r in best:
                    print(f"   {100 * int(r[ix['# Samples']]) / tot:5.1f} %  exec {int(r[ix['Instructions Executed']]) / 1e6:7.2f} M  "
                          f"thr {r[ix['Avg. Threads Executed']]:>4}  long_sb {r[ix['stall_long_sb']]:>6}  {r[ix['Source']].strip()[:70]}")


if __name__ == "__main__":
    main()
