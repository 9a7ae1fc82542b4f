"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line:
executed warp instructions and stall samples per line, per kernel."""
import csv
import sys
import collections

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur_file, cur_fn, hdr = None, None, None
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, ""]))
totals = collections.defaultdict(lambda: [0, 0])
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1]; continue
    if row[0] == "Function Name":
        cur_fn = row[1]; continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or len(row) < 8:
        continue
    try:
        line = row[0]
        src = row[1]
        inst = int(row[hdr.index("Instructions Executed")] or 0)
        samp = int(row[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    if row[2] != "-":     # SASS rows carry an address (or "..."); CUDA rows ("-") hold the per-line totals
        continue
    key = (cur_file.split("/")[-1], line)
    a = agg[cur_fn][key]
    a[0] += inst; a[1] += samp; a[2] = src.strip()[:90]
    totals[cur_fn][0] += inst; totals[cur_fn][1] += samp
for fn, lines in agg.items():
    ti, ts = totals[fn]
    print("=" * 110); print(fn, "inst", ti, "samples", ts)
    for (f, ln), (inst, samp, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% inst %5.1f%% stall  %-16s:%-4s %s" % (100.0 * inst / max(ti, 1), 100.0 * samp / max(ts, 1), f, ln, src))
