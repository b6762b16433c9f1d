#!/bin/bash
# usage: scripts/r02_n.sh N "name|args" "name|args" ...   (sharded bench variants at N ranks)
N=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%|*}; args=${spec#*|}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 20 --warmup 3 $args > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n${N}_$name.json"))
    print("$name", "ms", d["ms_per_step"], "G/s", round(d["value"]/1e9,2), "verified", d.get("verified"), {k:v["ms"] for k,v in d["stages"].items()}, d["config"].get("peer_wait_status"))
except Exception as e:
    print("$name failed", e)
PY
  grep -i "error\|Traceback" -A3 gpurun_out/bench_n${N}_$name.err | head -12
done
