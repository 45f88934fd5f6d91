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


def traffic_json(out_path, reps):
    """DRAM bytes (read + write) per launch of the captured kernels -> the JSON bench.py reads for `roofline.traffic`."""
    import json
    import os
    import re
    res = {"source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch (tools/gpu_ncu_r02.sh)", "kernels": []}
    for path in reps:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]

        def val(r, k):
            v = float(r[hdr.index(k)].replace(",", ""))
            u = units[hdr.index(k)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        tot, n = 0.0, 0
        for r in rows[2:]:
            by = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
            res["kernels"].append({"report": os.path.basename(path), "kernel": r[hdr.index("Kernel Name")].split("(")[0][-40:],
                                   "grid": r[hdr.index("launch__grid_size")], "dram_bytes": by,
                                   "time_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
                                   {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1)})
            tot += by
            n += 1
        m = re.search(r"_(gemm|k_attend|k_vocab_merge|k_beam_step)_b(\d+)", os.path.basename(path))
        if m and n:
            key = {"gemm": "gemm", "k_attend": "attend", "k_vocab_merge": "vocab_merge", "k_beam_step": "beam_step"}[m.group(1)] + "_bytes_per_launch_b" + m.group(2)
            res[key] = tot / n
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--traffic-json":
        traffic_json(sys.argv[2], sys.argv[3:])
    else:
        for p in sys.argv[1:]:
            (launch_list if p.endswith(".csv") else full_report)(p)
