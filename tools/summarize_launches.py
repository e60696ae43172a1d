#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv [--skip N] > profiles/rNN_launches_summary.txt

Per-launch times from ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md).
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(?:<[^(]{0,60})?)", name)
    return (m.group(1) if m else name)[:96]


def owner(name: str) -> str:
    if "hg::" in name:
        return "libhologan_b200 (hand-written sm_100a)"
    if any(t in name for t in ("cudnn", "cutlass", "nvjet", "implicit_convolve", "convolve_common", "gemv", "sgemm", "gemm")):
        return "cuDNN / cuBLAS"
    if "nccl" in name.lower():
        return "NCCL"
    return "torch (elementwise / reduce / copy / optimizer)"


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
            rows.append((int(r["ID"]), r["Kernel Name"], v))
    rows = [r for r in rows if r[0] >= skip]
    tot = sum(r[2] for r in rows)
    per, own = defaultdict(lambda: [0, 0.0]), defaultdict(lambda: [0, 0.0])
    for _, k, ns in rows:
        per[short(k)][0] += 1; per[short(k)][1] += ns
        own[owner(k)][0] += 1; own[owner(k)][1] += ns
    print(f"# {path}: {len(rows)} launches (skipped ids < {skip}), total {tot / 1e3:.1f} us (serialised, cold-cache)")
    print("\n## by owner")
    for k, (n, ns) in sorted(own.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / tot * 100:6.2f} %  {ns / 1e3:10.1f} us  x{n:5d}  {k}")
    print("\n## by kernel")
    for k, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{ns / tot * 100:6.2f} %  {ns / 1e3:10.1f} us  x{n:5d}  avg {ns / n / 1e3:8.1f} us  {k}")


if __name__ == "__main__":
    main()
