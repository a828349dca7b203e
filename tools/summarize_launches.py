"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv [bench.json] > profiles/rNN_launches.md

Per-launch times under ncu are cold-cache and serialised, so only each kernel's SHARE of the step is
comparable with the live CUDA-event numbers of bench.py (second argument, optional).
"""
import collections
import csv
import json
import re
import sys


def short(name):
    m = re.search(r"(\w+)(<[^(]*>)?\(", name)
    return m.group(1) if m else name


def main():
    path = sys.argv[1]
    bench = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else None
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            ns = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = short(r[ki])
        a = agg.setdefault(k, {"n": 0, "ns": 0.0, "grid": r[gi], "block": r[bi]})
        a["n"] += 1
        a["ns"] += ns
    total = sum(a["ns"] for a in agg.values())
    print("| kernel | launches | total ms | mean us | share under ncu | share live (bench.py) | grid | block |")
    print("|---|---|---|---|---|---|---|---|")
    live = (bench or {}).get("kernels", {})
    live_total = sum(v["ms_per_launch"] * v["launches_per_step"] for v in live.values()) or 1.0

    def live_share(k):
        cands = [v for name, v in live.items() if name == k or name.startswith(k.replace("_kernel", "")) and k.startswith("huffman_build")]
        if not cands:
            return ""
        return "%.1f%%" % (100 * sum(v["ms_per_launch"] * v["launches_per_step"] for v in cands) / live_total)

    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print("| %s | %d | %.3f | %.1f | %.1f%% | %s | %s | %s |" % (k, a["n"], a["ns"] / 1e6, a["ns"] / a["n"] / 1e3,
                                                                    100 * a["ns"] / total, live_share(k), a["grid"], a["block"]))
    print()
    print("total kernel time in the list: %.2f ms over %d launches" % (total / 1e6, sum(a["n"] for a in agg.values())))


if __name__ == "__main__":
    main()
