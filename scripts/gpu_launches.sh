#!/bin/bash
# ncu launch list of the bench step (cold cache, serialised): every kernel of
# this library, last full step printed.
mkdir -p gpurun_out
KREGEX='regex:FwdPool|FwdConcat|Radix|Compress|RowIds|Bwd'
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k "$KREGEX" -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches.csv', errors='ignore')))
i0=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
data=rows[i0+1:]
# one step = 12 launches (fwd, row ids, hist, 4 passes, 2 compress, 3 backward)
for r in data[-12:]:
    print(r[4][:60].ljust(60), r[8].ljust(14), r[-1])
PY
