"""Per-CUDA-source-line stall samples and executed instructions from an ncu report captured with
--import-source on (ncu -i REP --page source --print-source sass,cuda --csv | python tools/ncu_lines.py)."""
import csv
import sys


def main():
    rows = list(csv.reader(sys.stdin))
    cur_file, hdr, agg = None, None, {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) > 6 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        try:
            si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        except ValueError:
            si, ii = 4, hdr.index("Instructions Executed")
        if r[2] != "-":          # SASS rows repeat the totals of their line row
            continue
        key = (cur_file, int(r[0]))
        s = int(r[si]) if r[si].isdigit() else 0
        n = int(r[ii]) if r[ii].isdigit() else 0
        a = agg.setdefault(key, [0, 0, r[1].strip()[:100]])
        a[0] += s
        a[1] += n
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print(f"total samples {tot}, instructions {toti}")
    for (f, ln), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[1]) if len(sys.argv) > 1 else 30]:
        print(f"{100 * s / tot:5.1f}% samples {100 * n / toti:5.1f}% inst  {f}:{ln}: {src}")


if __name__ == "__main__":
    main()
