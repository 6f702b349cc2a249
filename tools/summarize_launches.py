"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name, grid, block) the number of
launches, mean duration and share of the total.  Usage: python tools/summarize_launches.py launches.csv [out.txt]"""
import collections
import csv
import sys


def summarize(path):
    rows = list(csv.reader(open(path)))
    hdr, agg, order = None, collections.OrderedDict(), []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = d["Kernel Name"]
        name = name.replace("void ", "").replace("lgae::", "")
        name = name.split("(")[0]
        key = (name, d["Grid Size"], d["Block Size"])
        agg.setdefault(key, []).append(v_us)
    total = sum(sum(v) for v in agg.values())
    lines = [f"{'launches':>8} {'mean us':>10} {'sum us':>10} {'share':>7}  kernel  grid  block"]
    for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{len(v):8d} {sum(v) / len(v):10.1f} {sum(v):10.1f} {100 * sum(v) / total:6.1f}%  {name}  {grid}  {block}")
    lines.append(f"total {total:.1f} us over {sum(len(v) for v in agg.values())} launches")
    return "\n".join(lines)


if __name__ == "__main__":
    out = summarize(sys.argv[1])
    print(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(out + "\n")
