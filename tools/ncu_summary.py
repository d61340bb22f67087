#!/usr/bin/env python
"""Summarise an ncu report (key metrics, opcode histogram per warp-point, top stall sites) -- run here, no GPU needed.
usage: tools/ncu_summary.py <report.ncu-rep> <points> [kernel-regex]"""
import collections, csv, io, re, subprocess, sys

rep, npts = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = dict(zip(hdr, vals))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]
for k in keys:
    for h in hdr:
        if h == k or h.startswith(k):
            print("%-75s %s %s" % (h, M[h], units[hdr.index(h)])); break
wp = npts / 32
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
try:
    print("warp instr / warp-point: %.1f" % (float(M["smsp__inst_executed.sum"].replace(",", "")) / wp))
    rd = float(M["dram__bytes_read.sum"].replace(",", "")) * SCALE[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(M["dram__bytes_write.sum"].replace(",", "")) * SCALE[units[hdr.index("dram__bytes_write.sum")]]
    print("dram bytes / launch: read %.4g write %.4g total %.4g   per point: read %.1f write %.1f total %.1f" %
          (rd, wr, rd + wr, rd / npts, wr / npts, (rd + wr) / npts))
except Exception as e:
    print("derived metrics unavailable:", e)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(hdr)}
cnt = collections.Counter(); tot = 0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; ss = collections.Counter()
for r in data:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    if not m: continue
    cnt[m.group(2).split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
    tot += int(r[ix["# Samples"]] or 0)
    for s in stalls: ss[s] += int(r[ix[s]] or 0)
print("opcodes per warp-point:", ", ".join("%s %.1f" % (o, v / wp) for o, v in cnt.most_common(24)))
print("stalls:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(tot, 1)) for s, v in ss.most_common(8)))
print("top stall sites:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print("  %6s %9s  %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:90]))
