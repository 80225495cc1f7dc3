#!/usr/bin/env python3
"""Extract the metrics the roofline argument rests on from a `ncu --set full` report into a small
text file that can live in git (the .ncu-rep files themselves are 10-25 MB each).

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01/name.txt ["free-text note"]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def raw(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = raw(rep, "raw")
    hdr, units = rows[0], rows[1]
    lines = [f"# {rep}", f"# {note}" if note else "#", "# ncu --set full --clock-control none (one launch, replayed)", ""]
    for r in rows[2:3]:
        lines.append(f"kernel: {r[hdr.index('Kernel Name')]}")
        for m in METRICS:
            if m in hdr:
                lines.append(f"{m:86s} {r[hdr.index(m)]:>20s} {units[hdr.index(m)]}")
    src = raw(rep, "source")
    for i, r in enumerate(src):
        if "Source" in r and "# Samples" in r:
            h = r
            si, ci, ei = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
            data, seen = [], set()
            for rr in src[i + 1:]:
                try:
                    key = (rr[0], rr[si])
                    if key in seen:
                        continue
                    seen.add(key)
                    data.append((float(rr[ci].replace(",", "")), float(rr[ei].replace(",", "")), rr[si]))
                except (ValueError, IndexError):
                    continue
            tot = sum(d[0] for d in data) or 1.0
            lines += ["", f"top warp-stall sample sites (of {int(tot)} samples, {len(data)} SASS instructions):"]
            for v, e, s in sorted(data, key=lambda t: -t[0])[:14]:
                lines.append(f"  {100 * v / tot:6.2f}%  {s.strip()[:100]}")
            break
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
