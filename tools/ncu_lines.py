"""Per-source-line stall profile from an ncu report: joins `ncu --page source --csv` (SASS view, one row per
instruction) with `nvdisasm -g` line info of the same kernel.
Usage: python tools/ncu_lines.py report.ncu-rep file.cubin <mangled-name-substring> [top]"""
import collections
import csv
import io
import re
import subprocess
import sys


def sass_lines(cubin, func):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    res, cur, active = [], None, False
    for ln in out:
        if ln.startswith(".text."):
            active = func in ln
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            res.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return res


def main():
    rep, cubin, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    data = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) == len(hdr)]
    sl = sass_lines(cubin, func)
    assert len(sl) == len(data), (len(sl), len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    total = 0
    for (off, loc, ins), d in zip(sl, data):
        n = int(d["# Samples"] or 0)
        total += n
        a = agg[loc]
        a["samples"] += n
        a["inst"] += int(d["Instructions Executed"] or 0)
        for s in stalls:
            a[s] += int(d[s] or 0)
    print(f"total samples {total}")
    tot_st = collections.Counter()
    for a in agg.values():
        for s in stalls:
            tot_st[s] += a[s]
    print("stall totals:", ", ".join(f"{k[6:]} {100 * v / total:.1f}%" for k, v in tot_st.most_common(9)))
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = ", ".join(f"{k[6:]} {v}" for k, v in sorted(((s, a[s]) for s in stalls), key=lambda kv: -kv[1])[:4] if v)
        print(f"{loc[0] if loc else '?'}:{loc[1] if loc else 0:5d}  {100 * a['samples'] / total:5.1f}%  inst {a['inst']:9d}  {st}")


if __name__ == "__main__":
    main()
