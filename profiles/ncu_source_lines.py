#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep captured with
--import-source on (needs -lineinfo).  python profiles/ncu_source_lines.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur_file = ""; agg = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "": continue
    i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    try:
        agg.append((int(r[i_s]), int(r[i_i]), cur_file, int(r[0]), r[1].strip()[:100]))
    except ValueError:
        pass
tot_s = sum(a[0] for a in agg) or 1; tot_i = sum(a[1] for a in agg) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
for s, i, f, ln, src in sorted(agg, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, f, ln, src))
