#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py gpurun_out/r01a/lk_track_smem_kernel.ncu-rep [top_n] [bucket_lines]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; per = collections.OrderedDict(); src = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    key = (cur_file, int(r[0]))
    src[key] = r[1]
    i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed")
    a = per.setdefault(key, [0, 0])
    try:
        a[0] += int(r[i_s]); a[1] += int(r[i_i])
    except ValueError:
        pass
tot = sum(a[0] for a in per.values()) or 1
print("total samples", tot)
for key, a in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%%  inst %9d  %s:%d  %s" % (100.0 * a[0] / tot, a[1], key[0], key[1], src[key].strip()[:110]))
