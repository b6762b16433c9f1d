#!/usr/bin/env python
"""Print the top stall lines (SASS) of an ncu report: python scripts/ncu_top.py rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; si = hdr.index('# Samples'); src = hdr.index('Source')
data = []
for k, r in enumerate(rows[2:]):
    try: data.append((int(r[si]), k, r[src].strip()))
    except Exception: pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
for s, k, t in sorted(data, reverse=True)[:n]:
    print('%6d %5.1f%% line %4d  %s' % (s, 100 * s / tot, k, t[:100]))
