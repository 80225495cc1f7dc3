#!/usr/bin/env python3
"""Write profiles/roofline_traffic.json from the committed `ncu --set full` summaries (profiles/summarize_ncu.py): DRAM
bytes per STEP = dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step, stamped with the commit and
the hash of the kernel sources (bench.py: kernel_source_hash) the captures were taken from.

    python profiles/stamp_traffic.py <dir with ncu_*.txt> <commit>
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
STEPS = {   # workload -> (summaries of the launches of one step, queries per step at the capture)
    "c2": (["ncu_c2_cubic.txt"], 1 << 20),
    "c3": (["ncu_c3_linear_pair.txt"], 1 << 24),
    "c4": (["ncu_c4_bilinear.txt"], 1 << 24),
    "c5a": (["ncu_c5a_bilinear_binned.txt", "ncu_c5a_bin_scatter.txt", "ncu_c5a_bin_totals.txt"], 1 << 28),
    "c5b": (["ncu_c5b_cubic.txt"], 1 << 28),
}


def dram_bytes(path):
    tot = 0.0
    for ln in open(path):
        m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", ln)
        if m:
            tot += float(m.group(2)) * UNIT[m.group(3)]
    return int(tot)


def main():
    d, commit = sys.argv[1], sys.argv[2]
    import bench
    out = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per STEP (all launches of the step) from `ncu --set full` "
                    "captures summarised under profiles/; each entry names the commit and the hash of the evaluation-kernel "
                    "sources of its workload (bench.py: workload_sources, kernel_source_hash) it was taken from -- "
                    "bench.py reports `current: false` when those sources have changed since; written by "
                    "profiles/stamp_traffic.py"}
    try:                                             # entries whose summaries are not in `d` keep their capture
        for k, v in json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).items():
            if k != "_note":
                out[k] = v
    except Exception:
        pass
    for wl, (files, nq) in STEPS.items():
        paths = [os.path.join(d, f) for f in files]
        if not all(os.path.exists(p) for p in paths):
            continue
        out[wl] = {"bytes": sum(dram_bytes(p) for p in paths), "per": f"step of {nq} queries", "queries": nq, "commit": commit,
                   "source_hash": bench.kernel_source_hash(bench.workload_sources(wl)),
                   "capture": " + ".join(os.path.relpath(p, ROOT) for p in paths)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
