"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
                ki, vi, ui = r.index("Kernel Name"), r.index("Metric Value"), r.index("Metric Unit")
            continue
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "us" else (v / 1e6 if r[ui] == "ns" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[:64]:64s} n={v[0]:5d} total={v[1]:10.3f} ms  {100 * v[1] / tot:5.1f}%")
    print(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
