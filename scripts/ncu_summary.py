"""Summarise ncu artefacts brought back in gpurun_out/ into profiles/ (text, committed).
usage: python scripts/ncu_summary.py <launches.csv> <report.ncu-rep> [...] > profiles/<name>.md"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"## launch list `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {t / tot:.1%} |")
    print()


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print(f"## `{path}` (ncu --set full --clock-control none)\n")
    kn = h.index("Kernel Name")
    for r in rows[2:]:
        print(f"**{r[kn][:120]}**\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in h:
                print(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |")
        st = []
        for i, name in enumerate(h):
            if "pcsamp_warps_issue_stalled" in name and "not_issued" not in name:
                try:
                    st.append((float(r[i].replace(",", "")), name.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1
        print("\nwarp-state samples: " + ", ".join(f"{n} {v / tot:.0%}" for v, n in sorted(st, reverse=True)[:8]) + "\n")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (launches if p.endswith(".csv") else report)(p)
