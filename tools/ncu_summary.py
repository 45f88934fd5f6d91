"""Summarise ncu captures for profiles/: (1) per-kernel totals and shares from a launch list CSV
(--metrics gpu__time_duration.sum), (2) key metrics of `--set full` .ncu-rep captures."""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct"]


def launch_list(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
            rows.append((r["Kernel Name"].split("(")[0].replace("void ", ""), v_us))
    tot = OrderedDict()
    for k, v in rows:
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(v[1] for v in tot.values())
    print(f"# launch list {path}: {len(rows)} launches, {total/1e3:.3f} ms total (cold-cache, serialised)")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {t:.1f} | {t/n:.2f} | {100*t/total:.1f}% |")


def full_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"\n# ncu --set full: {path}")
    for r in rows[2:]:
        print(f"- {r[hdr.index('Kernel Name')]} grid={r[hdr.index('launch__grid_size')]}")
        for k in KEYS:
            if k in hdr:
                print(f"    {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        (launch_list if p.endswith(".csv") else full_report)(p)
