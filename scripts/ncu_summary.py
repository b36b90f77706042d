"""Summarise ncu captures into the tracked text files under profiles/.

    python scripts/ncu_summary.py full  gpurun_out/x.ncu-rep  profiles/r01_x_full.md
    python scripts/ncu_summary.py list  gpurun_out/launches.csv profiles/r01_launches.md [skip_first_n]

`full` reads an `ncu --set full` report (through `ncu -i ... --page raw --csv`) and writes one row per
captured launch with the metrics the roofline needs.  `list` reads the csv of a
`--metrics gpu__time_duration.sum` pass and writes per-kernel totals and shares.
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL_COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%act"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_%el"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wf_%"),
    ("smsp__inst_executed_op_shared_atom.sum", "atoms_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def short(name):
    name = name.replace("void ", "").replace("eds::", "")
    return name.split("(")[0][:60]


def to_num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def scale(v, unit):
    """Return (value in base unit) for the units ncu prints."""
    mult = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
            "Tbyte": 1e12}
    return v * mult.get(unit, 1.0)


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full summary of `%s`" % rep, "",
             "Per captured launch (cold-cache, serialised by ncu; use for traffic/utilisation, not for timing claims).",
             "", "| # | kernel | " + " | ".join(c[1] for c in FULL_COLS) + " |",
             "|---|---|" + "---|" * len(FULL_COLS)]
    for n, r in enumerate(body):
        cells = []
        for m, label in FULL_COLS:
            if m not in idx:
                cells.append("-")
                continue
            v = to_num(r[idx[m]])
            u = units[idx[m]]
            if v is None:
                cells.append(r[idx[m]])
            elif label == "time":
                cells.append("%.1f us" % (scale(v, u) * 1e6))
            elif label in ("dram_rd", "dram_wr", "l2_bytes"):
                cells.append("%.2f MB" % (scale(v, u) / 1e6))
            elif label in ("regs", "grid", "block", "atoms_inst"):
                cells.append("%d" % v)
            else:
                cells.append("%.1f" % v)
        lines.append("| %d | %s | %s |" % (n, short(r[idx["Kernel Name"]]), " | ".join(cells)))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def launch_list(path, out, skip=0):
    raw = open(path).read().splitlines()
    start = next(i for i, l in enumerate(raw) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(raw[start:]))))
    rows = rows[skip:]
    tot = OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        t = scale(v, u)
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in tot.values())
    lines = ["# ncu launch list summary of `%s` (first %d launches skipped)" % (path, skip), "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none`: per-launch device time, cold-cache and",
             "serialised, so SHARES are meaningful, absolutes are not.  %d launches, %.3f ms in total." % (
                 len(rows), total * 1e3), "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.3f | %.1f%% |" % (k, n, t * 1e3, 100 * t / total))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3])
    else:
        launch_list(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
