"""Per-opcode executed counts and the top stall lines of one kernel from `ncu --page source --csv --print-source sass` output.
    python tools/ncu_source_top.py source.csv [section_index] [units]   (units: divide executed counts by this, e.g. tiles)"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    i0 = secs[sec]
    end = secs[sec + 1] if sec + 1 < len(secs) else len(rows)
    hdr = rows[i0 + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    ie, isam, isrc = ix["Instructions Executed"], ix["# Samples"], ix["Source"]
    data = [r for r in rows[i0 + 2:end] if len(r) > ie and r[ie].isdigit()]
    print(rows[i0][1][:100])
    tot, tots = sum(int(r[ie]) for r in data), sum(int(r[isam]) for r in data)
    print(f"executed {tot} ({tot / units:.1f} per unit), static {len(data)}, samples {tots}")
    c, s = collections.Counter(), collections.Counter()
    for r in data:
        op = [o for o in r[isrc].split() if not o.startswith("@")]
        op = op[0].split(".")[0] if op else "?"
        c[op] += int(r[ie])
        s[op] += int(r[isam])
    for op, n in c.most_common(22):
        print(f"  {op:10s} {n / units:10.1f}   samples {s[op]}")
    print("top stall lines (samples, executed, sass)")
    for r in sorted(data, key=lambda r: -int(r[isam]))[:22]:
        print(f"  {r[isam]:>6s} {r[ie]:>9s}  {r[isrc][:100]}")


if __name__ == "__main__":
    main()
