"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and total time per kernel.

usage: python profiles/summarize_launches.py gpurun_out/r1_launches.csv > profiles/r1_launch_list.txt
"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        rows = list(csv.reader(line for line in f if not line.startswith("==")))
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        name = r[ci["Kernel Name"]]
        short = name.split("(")[0][-70:]
        key = (short, r[ci["Block Size"]], r[ci["Grid Size"]])
        ns = float(r[ci["Metric Value"]].replace(",", ""))
        a = agg.setdefault(key, [0, 0.0, []])
        a[0] += 1
        a[1] += ns
        a[2].append(ns)
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time (ncu, serialised)")
    print(f"{'n':>5} {'total ms':>10} {'share':>7} {'avg us':>10}  kernel  [block] [grid]")
    for (short, blk, grd), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[0]:5d} {a[1] / 1e6:10.3f} {a[1] / tot * 100:6.1f}% {a[1] / a[0] / 1e3:10.1f}  {short}  [{blk}] [{grd}]")
    print("# per-launch durations (us) of the drcvar kernels, in launch order")
    for (short, blk, grd), a in agg.items():
        if "drcvar" in short:
            print(f"{short} [{grd}]: " + " ".join(f"{ns / 1e3:.1f}" for ns in a[2]))


if __name__ == "__main__":
    main(sys.argv[1])
