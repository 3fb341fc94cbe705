"""`ncu --set full` report -> markdown table of the metrics the roofline claims rest on (one row per captured launch).

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep "<command that was profiled>" > profiles/<name>.md

Needs the `ncu` CLI (no GPU): reads the report with `ncu -i ... --page raw --csv`.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    ("gpu__time_duration.sum", "time", 1.0),
    ("dram__bytes_read.sum", "DRAM read", 1.0),
    ("dram__bytes_write.sum", "DRAM write", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak", 1.0),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active", 1.0),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % busy", 1.0),
    ("smsp__inst_executed.sum", "warp instructions", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1.0),
    ("launch__registers_per_thread", "regs/thread", 1.0),
    ("launch__grid_size", "grid", 1.0),
    ("launch__block_size", "block", 1.0),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(v.replace(",", "")) * mult


def to_us(v, unit):
    mult = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
    return float(v.replace(",", "")) * mult


def main(path, cmd):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    print("# ncu --set full: `%s`\n" % cmd)
    print("Report `%s` read with `ncu -i ... --page raw --csv` (clock control none; one replayed launch per row: cold "
          "caches, serialised).  DRAM GB/s = (read + write) / time; the fraction is against the measured copy bandwidth "
          "in MEASURED_PEAKS.json (%s GB/s).\n" % (os.path.basename(path), peaks.get("hbm_gbs", "n/a")))
    names = ["kernel", "time us", "DRAM read MB", "DRAM write MB", "DRAM GB/s", "of measured HBM", "tensor pipe % active",
             "issue slots % busy", "warp instr", "regs", "grid x block", "L2 hit %"]
    print("| " + " | ".join(names) + " |")
    print("|" + "---|" * len(names))
    for r in body:
        if len(r) < len(hdr):
            continue
        def get(m):
            return (r[col[m]], units[col[m]]) if m in col else ("0", "")
        name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        t = to_us(*get("gpu__time_duration.sum"))
        rd, wr = to_bytes(*get("dram__bytes_read.sum")), to_bytes(*get("dram__bytes_write.sum"))
        gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else 0
        frac = "%.2f" % (gbs / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else "n/a"
        f = lambda m: get(m)[0].replace(",", "")
        print("| `%s` | %.1f | %.1f | %.1f | %.0f | %s | %.1f | %.1f | %s | %s | %s x %s | %.1f |" % (
            name[:70], t, rd / 1e6, wr / 1e6, gbs, frac, float(f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0),
            float(f("smsp__issue_active.avg.pct_of_peak_sustained_active") or 0), f("smsp__inst_executed.sum").split(".")[0],
            f("launch__registers_per_thread").split(".")[0], f("launch__grid_size").split(".")[0],
            f("launch__block_size").split(".")[0], float(f("lts__t_sector_hit_rate.pct") or 0)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
