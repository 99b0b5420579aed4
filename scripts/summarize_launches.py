"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python scripts/summarize_launches.py launches.csv [first_id last_id]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, vi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
    out = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi or not r[ii].isdigit():
            continue
        out.append((int(r[ii]), r[ki], float(r[vi].replace(",", ""))))
    return out


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name[:90]


def main():
    rows = load(sys.argv[1])
    if len(sys.argv) > 3:
        lo, hi = int(sys.argv[2]), int(sys.argv[3])
        rows = [r for r in rows if lo <= r[0] <= hi]
    agg = collections.OrderedDict()
    for _, k, v in rows:
        a = agg.setdefault(short(k), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {tot / 1e6:.3f} ms kernel time")
    print("| kernel | launches | total ms | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / a[0] / 1e3:.2f} | {100 * a[1] / tot:.2f}% |")


if __name__ == "__main__":
    main()
