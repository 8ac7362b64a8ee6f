"""Turns `ncu -i X.ncu-rep --page raw --csv` output into the two artefacts kept under profiles/:
    python tools/ncu_summary.py raw.csv --ops a,b,c --title "..." --txt profiles/rN_ncu_summary.txt --traffic profiles/rN_traffic.json
    python tools/ncu_summary.py launches.csv --launches --title "..." --txt profiles/rN_ncu_launches_summary.txt
--ops names the plan launches in capture order (tools/profile_ops.py --ops ...); --launches aggregates a
`--metrics gpu__time_duration.sum` launch list by kernel name."""
import argparse
import collections
import csv
import json
import re

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--ops", default="")
    ap.add_argument("--launches", action="store_true")
    ap.add_argument("--title", default="")
    ap.add_argument("--txt")
    ap.add_argument("--traffic")
    a = ap.parse_args()
    rows = [r for r in csv.reader(open(a.csv)) if r]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    iname = hdr.index("Kernel Name")
    out = [a.title] if a.title else []
    if a.launches:
        # `ncu --csv --log-file` long format: one row per (launch, metric)
        im, iu, iv = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        agg = collections.OrderedDict()
        for r in rows[h + 1:]:
            if len(r) <= iv or r[im] != "gpu__time_duration.sum":
                continue
            name = re.sub(r"\(.*", "", r[iname]).replace("void ", "").replace("ach::", "")
            t = num(r[iv])
            if t is None:
                continue
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[iu], 1e-6)
            n, s = agg.get(name, (0, 0.0))
            agg[name] = (n + 1, s + t * scale)
        total = sum(s for _, s in agg.values())
        out.append(f"total {total:.3f} ms over {sum(n for n, _ in agg.values())} launches")
        for name, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append(f"{name:62s} launches {n:5d}  time {s:9.3f} ms  share {100 * s / total:5.1f}%")
    else:
        units, data = rows[h + 1], rows[h + 2:]
        ops = a.ops.split(",") if a.ops else [f"launch{i}" for i in range(len(data))]
        assert len(ops) == len(data), (len(ops), len(data))
        out.append("launch: " + " | ".join(ops))
        out.append("kernel: " + " | ".join(r[iname][:48] for r in data))
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"{k} [{units[i]}]: " + " | ".join(r[i] for r in data))
        if a.traffic:
            ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
            mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tr = {}
            for op, r in zip(ops, data):
                tr[op] = {"dram_bytes": num(r[ir]) * mul[units[ir]] + num(r[iw]) * mul[units[iw]],
                          "duration_us": num(r[it]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[it]], "kernel": r[iname][:60]}
            with open(a.traffic, "w") as f:
                json.dump(tr, f, indent=1)
    text = "\n".join(out) + "\n"
    if a.txt:
        with open(a.txt, "w") as f:
            f.write(text)
    print(text)


if __name__ == "__main__":
    main()
