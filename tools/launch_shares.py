#!/usr/bin/env python
"""Per-kernel launch counts, total device time and share from an ncu `--metrics gpu__time_duration.sum --csv` launch list.
usage: tools/launch_shares.py <launches.csv>"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    n = r[ix["Kernel Name"]].split("(")[0][-60:]; v = float(r[ix["Metric Value"]].replace(",", ""))
    t[n][0] += 1; t[n][1] += v
tot = sum(v[1] for v in t.values())
print("# cold-cache, serialised launches under ncu: compare SHARES, not absolutes")
for n, (c, v) in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print("%-62s launches %3d  total %10.1f us  share %5.1f%%" % (n, c, v / 1e3, 100 * v / tot))
