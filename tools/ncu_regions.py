"""Buckets executed instructions of an ncu source export by enclosing source function."""
import collections
import csv
import os
import re
import sys

path = sys.argv[1]
cur_file = cur_fn = hdr = None
data = collections.defaultdict(lambda: collections.defaultdict(int))
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        cur_fn = row[1]; continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or len(row) < 8 or row[2] != "-":
        continue
    try:
        inst = int(row[hdr.index("Instructions Executed")] or 0); line = int(row[0])
    except ValueError:
        continue
    data[cur_fn][(cur_file, line)] += inst
src_cache = {}


def func_of(file, line):
    for root in ("/root/repo/tightly-coupled-sfm_b200/csrc", "/usr/local/cuda/include/crt", "/usr/local/cuda/include"):
        p = os.path.join(root, file)
        if os.path.isfile(p):
            if p not in src_cache:
                src_cache[p] = open(p, errors="replace").read().split("\n")
            lines = src_cache[p]
            for i in range(min(line, len(lines)) - 1, -1, -1):
                m = re.match(r"^\s*(?:__device__|__global__|static|inline|template|extern|__SM|__DEVICE|pair_|ssim_|warp_).*?(\w+)\s*\(", lines[i])
                if m and not lines[i].strip().startswith("//"):
                    return m.group(1)
            return file
    return file


for fn, lines in data.items():
    tot = sum(lines.values())
    agg = collections.defaultdict(int)
    for (f, ln), inst in lines.items():
        agg[(f, func_of(f, ln))] += inst
    print("=" * 90); print(fn, tot)
    for (f, name), inst in sorted(agg.items(), key=lambda kv: -kv[1])[:25]:
        print("%5.1f%%  %-22s %s" % (100.0 * inst / tot, f, name))
