#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table
(kernel, launches, total us, share). Usage: summarize_launches.py launches.csv [title] > profiles/x.md"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(x["Metric Value"].replace(",", ""))
        unit = x["Metric Unit"]
        v = v / 1e3 if unit in ("nsecond", "ns") else v * 1e3 if unit in ("msecond", "ms") else v
        a = agg.setdefault(x["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print("# %s\n" % title)
    print("%d launches, %.1f us total device time (ncu per-launch times are cold-cache and serialised: compare "
          "shares, not absolutes).\n" % (n, tot))
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k[:110].replace("|", "/"), a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


if __name__ == "__main__":
    main()
