#!/bin/bash
mkdir -p gpurun_out
KREGEX='regex:FwdPool|FwdConcat|Radix|Compress|RowIds|BwdSeg|BwdFix'
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k "$KREGEX" -s 36 -c 12 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches.csv')))
i0=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
for r in rows[i0+1:]:
    print(r[4][:60].ljust(60), r[8].ljust(14), r[-1])
PY
