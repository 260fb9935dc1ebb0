#!/usr/bin/env python
"""Join an `ncu --page source --csv` dump (per SASS instruction) with nvdisasm's line info of the same cubin:
per source line: instructions executed, stall samples.  usage: ncu_lines.py <ncu_source.csv> <nvdisasm_line_info.sass> <kernel-substring> [top]"""
import csv, re, sys
from collections import defaultdict

src_csv, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: instruction order -> source line
lines, cur, infn = [], None, False
for ln in open(sass, errors="replace"):
    if ln.startswith(".text."):
        infn = kname in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        inl = " (inlined)" if "inlined at" in ln else ""
        cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append((cur, ln.strip()))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
body = rows[2:]
print("sass instructions: nvdisasm %d, ncu %d" % (len(lines), len(body)))
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot_inst = tot_samp = 0
for i, r in enumerate(body):
    line = lines[i][0] if i < len(lines) else "?"
    inst = int(float(r[ix["Instructions Executed"]] or 0)); samp = int(float(r[ix["# Samples"]] or 0))
    a = agg[line]; a[0] += inst; a[1] += samp
    for c in stall_cols:
        v = int(float(r[ix[c]] or 0))
        if v: a[2][c] += v
    tot_inst += inst; tot_samp += samp
print("total warp instructions %d, samples %d" % (tot_inst, tot_samp))
for line, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    s = ", ".join("%s %d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print("%-28s inst %6.2f%%  samples %6.2f%%  [%s]" % (line, 100.0 * inst / tot_inst, 100.0 * samp / max(1, tot_samp), s))

# optional: aggregate by named line ranges of raster_view.cuh  (env NCU_RANGES="name:lo-hi,...")
import os
rng = os.environ.get("NCU_RANGES")
if rng:
    print("--- by range")
    for spec in rng.split(","):
        name, r = spec.split(":"); lo, hi = map(int, r.split("-"))
        inst = samp = 0
        for line, (i2, s2, st) in agg.items():
            if line and line.startswith("raster_view.cuh:"):
                n = int(line.split(":")[1])
                if lo <= n <= hi: inst += i2; samp += s2
        print("%-18s inst %6.2f%%  samples %6.2f%%" % (name, 100.0 * inst / tot_inst, 100.0 * samp / max(1, tot_samp)))
    other_i = sum(v[0] for k, v in agg.items() if not (k and k.startswith("raster_view.cuh:")))
    other_s = sum(v[1] for k, v in agg.items() if not (k and k.startswith("raster_view.cuh:")))
    print("%-18s inst %6.2f%%  samples %6.2f%%" % ("other files", 100.0 * other_i / tot_inst, 100.0 * other_s / max(1, tot_samp)))
