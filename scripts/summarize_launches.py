"""ncu launch list (gpu__time_duration.sum CSV) -> per-kernel summary table (markdown on stdout).

    python scripts/summarize_launches.py gpurun_out/launches.csv "<command that was profiled>" > profiles/<name>.md
"""
import collections
import csv
import sys


def main(path, cmd):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"].startswith("us"):
            ns *= 1e3
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if k.startswith("niw::") or not k.startswith(("void ", "at_cuda", "native")))
    print("# ncu launch list: `%s`\n" % cmd)
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare "
          "SHARES, not absolutes).  %d launches, %.1f us total, %.1f %% in this repo's kernels.\n"
          % (len(rows), total / 1e3, 100 * ours / total))
    print("| kernel | launches | grid | block | avg us | total us | share |")
    print("|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda t: -t[1][1]):
        print("| `%s` | %d | %s | %s | %.2f | %.1f | %.1f %% |" % (k[:90], a[0], a[2], a[3], a[1] / a[0] / 1e3, a[1] / 1e3,
                                                                  100 * a[1] / total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
