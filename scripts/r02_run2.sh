#!/bin/bash
# Run 2: full GPU test suite on the warp walker + A/B of backward variants.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$B > gpurun_out/bench_warp1.json 2> gpurun_out/bench_warp1.err; echo "warp=1 rc=$?"
CUEMBED_BWD_WARP=0 $B > gpurun_out/bench_warp0.json 2>/dev/null; echo "warp=0 rc=$?"
CUEMBED_B200_LIB=$PWD/cuembed_b200/lib/libcuembed_b200_minb5.so $B > gpurun_out/bench_minb5.json 2>/dev/null; echo "minb5 rc=$?"
CUEMBED_B200_LIB=$PWD/cuembed_b200/lib/libcuembed_b200_minb8.so $B > gpurun_out/bench_minb8.json 2>/dev/null; echo "minb8 rc=$?"
for f in warp1 warp0 minb5 minb8; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$f.json"))
    print("$f", d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items() if "ms" in v})
except Exception as e:
    print("$f", "failed", e)
PY
done
bash scripts/gpu_launches.sh
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:BwdWarpKernel -s 1 -c 1 -f -o gpurun_out/prof_BwdWarpKernel \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_BwdWarpKernel.log 2>&1
echo "ncu rc=$?"
