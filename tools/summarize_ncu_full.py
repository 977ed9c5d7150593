#!/usr/bin/env python
"""Summarise `ncu --set full` captures (.ncu-rep) into a markdown table: per launch duration, DRAM bytes and throughput,
pipe / issue / occupancy figures, and the warp-state sample breakdown (pc sampling). Needs `ncu` on PATH (reads the
report with `ncu -i ... --page raw --csv`; no GPU).
    python tools/summarize_ncu_full.py title a.ncu-rep [b.ncu-rep ...] > profiles/x.md"""
import csv
import io
import subprocess
import sys

COLS = [("time us", "gpu__time_duration.sum"), ("dram read MB", "dram__bytes_read.sum"), ("dram write MB", "dram__bytes_write.sum"),
        ("dram TB/s", "dram__bytes.sum.per_second"), ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("warp inst", "smsp__inst_executed.sum"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
        ("block", "launch__block_size"), ("dyn smem KB", "launch__shared_mem_per_block_dynamic")]
SCALE = {"nsecond": ("us", 1e-3), "usecond": ("us", 1.0), "msecond": ("us", 1e3), "byte": ("MB", 1e-6), "Kbyte": ("MB", 1e-3),
         "Mbyte": ("MB", 1.0), "Gbyte": ("MB", 1e3), "Gbyte/s": ("TB/s", 1e-3), "Tbyte/s": ("TB/s", 1.0), "Mbyte/s": ("TB/s", 1e-6)}


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    title, paths = sys.argv[1], sys.argv[2:]
    print("# %s\n" % title)
    print("`ncu --set full --clock-control none --import-source on`, one row per captured launch (cold-cache, serialised).\n")
    print("| kernel (report) | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    stalls = []
    for p in paths:
        hdr, units, rows = raw(p)
        for r in rows:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
            cells = []
            for _, m in COLS:
                if m not in hdr:
                    cells.append("-")
                    continue
                i = hdr.index(m)
                v, u = num(r[i]), units[i]
                if v is None:
                    cells.append(r[i])
                    continue
                if u in SCALE:
                    v *= SCALE[u][1]
                if m == "launch__shared_mem_per_block_dynamic" and u == "byte":
                    v *= 1e3   # (byte -> MB above) -> KB
                cells.append(("%.0f" % v) if abs(v) >= 1000 else ("%.2f" % v if abs(v) < 10 else "%.1f" % v))
            print("| `%s` (%s) | %s |" % (name, p.split("/")[-1], " | ".join(cells)))
            pre = "smsp__pcsamp_warps_issue_stalled_"
            samp = {h[len(pre):]: num(r[i]) for i, h in enumerate(hdr) if h.startswith(pre) and not h.endswith("_not_issued")}
            tot = sum(v for v in samp.values() if v) or 1.0
            top = sorted(((v / tot, k) for k, v in samp.items() if v), reverse=True)[:6]
            stalls.append((name, p.split("/")[-1], top))
    print("\nWarp-state samples (share of all pc samples of the launch, top 6):\n")
    seen = set()
    for name, rep, top in stalls:
        if (name, rep) in seen:
            continue
        seen.add((name, rep))
        print("* `%s` (%s): %s" % (name, rep, ", ".join("%s %.0f %%" % (k, 100 * f) for f, k in top)))


if __name__ == "__main__":
    main()
