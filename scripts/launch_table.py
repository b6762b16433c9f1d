"""Prints kernel name / grid / duration from an ncu --csv launch list."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
pat = sys.argv[2].split(",") if len(sys.argv) > 2 else None
hdr = None
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        n = d["Kernel Name"]
        if pat is None or any(p in n for p in pat):
            print(n.replace("cuembed_b200::", "").replace("void ", "")[:60], d.get("Grid Size"), d["Metric Value"], d["Metric Unit"])
