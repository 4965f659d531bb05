#!/usr/bin/env python
"""Condenses ncu reports (gpurun_out/*.ncu-rep) into small text summaries under profiles/.

usage: tools/summarize_ncu.py <report.ncu-rep> <profiles/out.txt>
       tools/summarize_ncu.py --launches <launches.csv> <profiles/out.txt>
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__inst_executed.sum",
]


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {path}\n")
        for r in rows[2:]:
            f.write(f"\n== {r[idx['Kernel Name']]}  (id {r[idx['ID']]})\n")
            for k in KEEP:
                if k in idx:
                    f.write(f"{k:75s} {r[idx[k]]:>16s} {units[idx[k]]}\n")
            try:   # ncu scales every column's unit on its own (read in Gbyte, write in Mbyte): convert before adding
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_read.sum"]]]
                wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_write.sum"]]]
                f.write(f"{'traffic = dram read + write':75s} {(rd + wr) / 1e6:16.3f} Mbyte\n")
            except Exception:
                pass


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(lambda: [0, 0.0])
    order = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        if name not in agg:
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += ns
    total = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list of {path}\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        f.write(f"# {'kernel':88s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for name in sorted(order, key=lambda n: -agg[n][1]):
            n, ns = agg[name]
            f.write(f"{name[:90]:90s} {n:8d} {ns / 1e3:12.1f} {ns / 1e3 / n:10.1f} {100 * ns / total:6.1f}%\n")
        f.write(f"# total {total / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches\n")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[1], sys.argv[2])
