#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of an ncu report captured with --import-source on
(kernel compiled with -lineinfo) -- run here, no GPU needed.  usage: tools/ncu_lines.py <report.ncu-rep> [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
cur = None; agg = collections.defaultdict(lambda: [0, 0, ""])
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": iex = r.index("Instructions Executed"); ism = r.index("Warp Stall Sampling (All Samples)"); continue
    if r[2] == "-":
        try: ln = int(r[0]); ex = int(r[iex]); sm = int(r[ism])
        except ValueError: continue
        a = agg[(cur, ln)]; a[0] += ex; a[1] += sm; a[2] = r[1]
tex = sum(v[0] for v in agg.values()); tsm = sum(v[1] for v in agg.values())
print("executed warp instructions %d, stall samples %d" % (tex, tsm))
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-18s %4d %5.1f%% instr %5.1f%% samples  %s" % (f, ln, 100 * v[0] / tex, 100 * v[1] / tsm, v[2].strip()[:100]))
