"""Aggregate an `ncu --page source --csv` dump (SASS view) by opcode: executed warp-instructions and stall samples."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ops = collections.defaultdict(lambda: [0, 0])
stalls = collections.Counter()
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
    if not m: continue
    op = m.group(2).split(".")[0]
    ie = int(float(r[ci["Instructions Executed"]] or 0)); s = int(float(r[ci["# Samples"]] or 0))
    ops[op][0] += ie; ops[op][1] += s; tot_i += ie; tot_s += s
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h] += int(float(r[ci[h]] or 0))
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for op, (ie, s) in sorted(ops.items(), key=lambda t: -t[1][0])[:28]:
    print("%-12s inst %10d (%5.1f%%)  samples %7d (%5.1f%%)" % (op, ie, 100 * ie / tot_i, s, 100 * s / max(tot_s, 1)))
print("stall reasons:", [(k, v) for k, v in stalls.most_common(8)])
