#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` output (one `--set full` capture) into a per-launch table of
the metrics that decide the roofline (duration, DRAM bytes, L2 / TMA traffic, pipe utilisation, occupancy).

    python tools/summarize_ncu_full.py gpurun_out/rNN_full_rotate_raw.csv [more.csv ...] > profiles/rNN_ncu_full_summary.txt
    python tools/summarize_ncu_full.py --json key=substring[#occurrence] ... file.csv   # dram bytes per launch as JSON

Numbers are from a profiled (replayed, serialised) run: use them for TRAFFIC and pipe shares, never for speed.
"""
import csv
import json
import sys

COLS = [
    ("us", "gpu__time_duration.sum", 1.0),
    ("grid", "launch__grid_size", 1.0),
    ("block", "launch__block_size", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("dram_rd_MB", "dram__bytes_read.sum", 1.0),
    ("dram_wr_MB", "dram__bytes_write.sum", 1.0),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("l2_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("l1tex_%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("tma_ld_GB", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", 1.0),
    ("mem_tensor_%", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    ("issue_%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("occ_%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0),
]


def to_bytes(val: str, unit: str) -> float:
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    try:
        return float(val.replace(",", "")) * mult
    except ValueError:
        return float("nan")


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"name": r[hdr.index("Kernel Name")]}
        for tag, col, _ in COLS:
            if col in hdr:
                i = hdr.index(col)
                d[tag] = (r[i], units[i])
        out.append(d)
    return out


def main():
    args = sys.argv[1:]
    if args and args[0] == "--json":
        keys = [a for a in args[1:] if "=" in a]
        files = [a for a in args[1:] if "=" not in a]
        launches = [l for f in files for l in load(f)]
        res = {}
        for k in keys:
            key, pat = k.split("=", 1)
            occ = 0
            if "#" in pat:
                pat, occ = pat.rsplit("#", 1)
                occ = int(occ)
            hits = [l for l in launches if pat in l["name"]]
            if len(hits) > occ:
                l = hits[occ]
                rd, wr = to_bytes(*l["dram_rd_MB"]), to_bytes(*l["dram_wr_MB"])
                res[key] = {"kernel": l["name"][:120], "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                            "profiled_us": float(l["us"][0])}
        print(json.dumps(res, indent=1))
        return
    for f in args:
        print(f"## {f}")
        print(f"{'kernel':58s} " + " ".join(f"{t:>11s}" for t, _, _ in COLS))
        for l in load(f):
            cells = []
            for tag, _, _ in COLS:
                if tag not in l:
                    cells.append(f"{'-':>11s}")
                    continue
                v, u = l[tag]
                if tag.endswith("_MB"):
                    cells.append(f"{to_bytes(v, u) / 1e6:11.2f}")
                elif tag.endswith("_GB"):
                    cells.append(f"{to_bytes(v, u) / 1e9:11.3f}")
                else:
                    try:
                        cells.append(f"{float(v.replace(',', '')):11.1f}")
                    except ValueError:
                        cells.append(f"{v[:11]:>11s}")
            print(f"{l['name'][:58]:58s} " + " ".join(cells))
        print()


if __name__ == "__main__":
    main()
