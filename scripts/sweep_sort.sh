#!/bin/bash
python scripts/bench_transpose.py
CUEMBED_SORT_DEBUG=1 python scripts/bench_transpose.py
CUEMBED_SORT_ITEMS=16 python scripts/bench_transpose.py
python scripts/bench_transpose.py 4194304 10000000 i64
python scripts/bench_transpose.py 16777216 400000000
